// The 3-channel ends of the wgancls path as direct kernels (no patch matrix in HBM).
//
// t2i_deconv_img: the TRANSPOSE of a 4x4 / stride-2 SAME convolution whose image side has 3 channels:
//     out[n, 2p-1+kh, 2q-1+kw, c] += sum_ci a[n, p, q, ci] * W[(kh*4+kw)*3 + c][ci]
//   * the input gradient of d_net's first conv (model.py:135 through tf.gradients at :63): a = gradient at the
//     conv output [N,32,32,128], W = that conv's weights, out = dD/d image [N,64,64,3] (fp32);
//   * g_net's last transposed conv (model.py:218), optionally fused with the 3 -> 3 channel 3x3 conv and tanh
//     behind it (:219-221): a = h5 [N,32,32,128], out = u4 (kept for the backward pass), img = tanh(conv3x3(u4)).
//   One persistent CTA sweeps an image top to bottom in tiles of 128 patch pixels: TMA stages the activation
//   tile (K-major, 128B swizzle), one thread issues tcgen05.mma with N = 48 columns (16 taps x 3 channels) into a
//   double-buffered TMEM accumulator, four epilogue warps move the [128 x 48] patch contributions to shared
//   memory and gather every output pixel from the (at most) four patches that reach it -- the overlap-add of the
//   transposed conv happens on chip.  The previous tile's last patch row is carried in a ring slot, so that image
//   rows on a tile boundary are complete without atomics.  The weights (<= 32 KB) stay resident in shared memory.
//   HBM traffic = the activation tensor once + the image once (the patch-matrix form wrote and re-read
//   [pixels x 64] bf16 in between and ran a separate scatter kernel).
//
// t2i_dense_f32: the conditioning head (mean | log_sigma = lrelu(cond @ W + b), model.py:113-114) in fp32 SIMT.
//   log_sigma feeds exp() (model.py:121): a bf16 rounding of it is the largest single contributor to the generator's
//   forward error (tools/bf16_error_budget.py), and the 256 x 256 x 1024 product is far too small for a tensor tile.
#include "host_util.h"
#include "img_patch.cuh"
#include "planes.cuh"
#include "ptx.cuh"

namespace t2i {

constexpr int kDiThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kDiEpi = 256;
constexpr int kDiStages = 3;           // K <= 256: a tile is at most four chunks; two CTAs share an SM (see the host side)
constexpr int kDiABytes = 128 * 64 * 2;  // 128 patch pixels x 64 channels
constexpr int kDiBChunk = 64 * 64 * 2;   // 64 x 64 weight block
constexpr int kDiSPitch = 52;            // floats per patch-contribution row (48 used; 16-byte stores conflict-free)
constexpr int kDiN = 48;

struct alignas(64) DeconvImgParams {
    CUtensorMap a_map;
    CUtensorMap b_map;
    int N, P, Q, bp, tiles_per_img;
    int lg_q, urows, slots;
    int k_chunks, last_k_steps, n_pass, np, b_kn;
    const float* bias3;
    float* out;          // fp32 NHWC [N][2P][2Q][3]
    const float* w9;     // 3x3 conv HWIO [3][3][3][3] or NULL
    const float* b9;
    float* img;          // tanh(conv3x3(out) + b9) when w9 != NULL
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tma_load_3d(const void* tmap, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__global__ void __launch_bounds__(kDiThreads, 2) deconv_img_kernel(const __grid_constant__ DeconvImgParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int np = prm.np, kch = prm.k_chunks;
    const int H = 2 * prm.P, W = 2 * prm.Q;
    // rings: patch rows -- the bp rows of this tile + the carried one = bp + 1 slots, indexed by (row % slots);
    // image rows for the fused 3x3 conv -- 2*bp + 3 live rows fit in prm.urows (a power of two), (row & (urows - 1))
    const int slots = prm.slots;
    const int urows = prm.urows;
    const int lg_w = prm.lg_q + 1;
    // carve-up: [A stages][B: np x k_chunks blocks][S: slots x Q x 52 fp32][U: urows x W*3 fp32][w9 81 + b9 3 + bias 3][barriers]
    uint8_t* s_a = smem;
    uint8_t* s_b = s_a + kDiStages * kDiABytes;
    float* s_s = reinterpret_cast<float*>(s_b + np * kch * kDiBChunk);
    float* s_u = s_s + slots * prm.Q * kDiSPitch;
    float* s_w = s_u + (prm.w9 != nullptr ? urows * W * 3 : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_w + 96);
    uint64_t* empty_bar = full_bar + kDiStages;
    uint64_t* tmem_full = empty_bar + kDiStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* b_bar = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&prm.a_map);
        tma_prefetch_desc(&prm.b_map);
        for (int i = 0; i < kDiStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kDiEpi / 32);
        }
        mbar_init(b_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const int total_tiles = prm.tiles_per_img;    // per image
    if (warp == 0) {
        if (elect_one()) {
            // weights: resident for the whole kernel
            mbar_arrive_expect_tx(b_bar, static_cast<uint32_t>(np * kch * kDiBChunk));
            for (int pl = 0; pl < np; ++pl)
                for (int kc = 0; kc < kch; ++kc) {
                    uint8_t* dst = s_b + (pl * kch + kc) * kDiBChunk;
                    if (prm.b_kn) tma_load_3d(&prm.b_map, b_bar, dst, 0, kc * 64, pl);      // [K rows][64 columns]
                    else tma_load_3d(&prm.b_map, b_bar, dst, kc * 64, 0, pl);               // [64 rows][K columns]
                }
            int stage = 0;
            uint32_t phase = 0;
            for (int n = blockIdx.x; n < prm.N; n += gridDim.x)
                for (int t = 0; t < total_tiles; ++t)
                    for (int pass = 0; pass < prm.n_pass; ++pass) {
                        const int pa = (pass == 1) ? 1 : 0;
                        for (int kc = 0; kc < kch; ++kc) {
                            mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                            mbar_arrive_expect_tx(&full_bar[stage], kDiABytes);
                            tma_load_5d(&prm.a_map, &full_bar[stage], s_a + stage * kDiABytes, kc * 64, 0, t * prm.bp, n, pa);
                            if (++stage == kDiStages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, kDiN, 0, prm.b_kn ? 1 : 0);
            mbar_wait(b_bar, 0, 600);
            tc_fence_after();
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int n = blockIdx.x; n < prm.N; n += gridDim.x)
                for (int t = 0; t < total_tiles; ++t, ++it) {
                    const int acc = it & 1;
                    mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, 300 + acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 64;
                    uint32_t first = 0;
                    for (int pass = 0; pass < prm.n_pass; ++pass) {
                        const int pb = (pass == 2) ? 1 : 0;
                        for (int kc = 0; kc < kch; ++kc) {
                            mbar_wait(&full_bar[stage], phase, 200 + stage);
                            tc_fence_after();
                            const int k_steps = (kc == kch - 1) ? prm.last_k_steps : 4;
                            const uint64_t da = make_sw128_desc(smem_u32(s_a + stage * kDiABytes), 0, 1024);
                            const uint32_t sb = smem_u32(s_b + (pb * kch + kc) * kDiBChunk);
                            const uint64_t db = prm.b_kn ? make_sw128_desc(sb, 8192, 1024) : make_sw128_desc(sb, 0, 1024);
                            for (int k = 0; k < k_steps; ++k) {
                                umma_bf16(d_tmem, da + 2 * k, db + (prm.b_kn ? 128 * k : 2 * k), idesc, first);
                                first = 1;
                            }
                            umma_commit(&empty_bar[stage]);
                            if (pass == prm.n_pass - 1 && kc == kch - 1) umma_commit(&tmem_full[acc]);
                            if (++stage == kDiStages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
        }
    } else {
        // ------------------------------------------------ epilogue: patch contributions -> image rows (256 threads)
        // TMEM -> shared memory: warp w reads the 32 accumulator rows of lane quarter (w & 3); the two warps of a quarter
        // split the 48 columns (32 + 16).  Gather / conv: thread t owns image column (t & (W - 1)) and walks rows, so the
        // column-side indices (which two patch columns, which kw) are loop invariants and nothing divides.
        const int quarter = warp & 3;
        const int part = (warp - 2) >> 2;               // 0: columns 0..31, 1: columns 32..47
        const int row = quarter * 32 + lane;            // TMEM lane = patch pixel inside the tile
        const int et = threadIdx.x - 64;
        const int pl = row >> prm.lg_q, q = row & (prm.Q - 1);
        const bool fuse = prm.w9 != nullptr;
        if (et < 81) s_w[et] = fuse ? prm.w9[et] : 0.f;
        else if (et < 84) s_w[et] = fuse ? prm.b9[et - 81] : 0.f;
        else if (et < 87) s_w[et] = prm.bias3 != nullptr ? prm.bias3[et - 84] : 0.f;
        const int ow = et & (W - 1);
        const int row_step = kDiEpi >> lg_w;            // image rows covered per sweep of the 256 threads (>= 1: W <= 256)
        const int row_first = et >> lg_w;
        const int kw0 = (ow + 1) & 1;
        const int q_a = (ow + 1 - kw0) >> 1, q_b = q_a - 1;           // patch columns reaching this pixel (kw0, kw0 + 2)
        const bool qa_ok = q_a < prm.Q, qb_ok = q_b >= 0;
        int it = 0;
        for (int n = blockIdx.x; n < prm.N; n += gridDim.x)
            for (int t = 0; t < total_tiles; ++t, ++it) {
                const int acc = it & 1;
                const int p0 = t * prm.bp;
                mbar_wait(&tmem_full[acc], (it >> 1) & 1, 400 + acc);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
                float4* dst = reinterpret_cast<float4*>(s_s + ((((p0 + pl) % slots) << prm.lg_q) + q) * kDiSPitch);
                if (part == 0) {
                    uint32_t r0[32];
                    tmem_ld_32x32(taddr, r0);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    named_bar(1, kDiEpi);                         // the previous tile's gather / conv is finished
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]),
                                             __uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3]));
                } else {
                    uint32_t r1[16];
                    tmem_ld_32x16(taddr + 32, r1);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    named_bar(1, kDiEpi);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dst[8 + j] = make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]),
                                                 __uint_as_float(r1[4 * j + 2]), __uint_as_float(r1[4 * j + 3]));
                }
                named_bar(2, kDiEpi);
                const bool last = (t == total_tiles - 1);
                const int oh_lo = (p0 == 0) ? 0 : 2 * p0 - 1;
                const int oh_hi = last ? H - 1 : 2 * p0 + 2 * prm.bp - 2;       // inclusive
                const float bz0 = s_w[84], bz1 = s_w[85], bz2 = s_w[86];
                float* obase = prm.out + (static_cast<long long>(n) * H * W + ow) * 3;
                for (int oh = oh_lo + row_first; oh <= oh_hi; oh += row_step) {
                    const int kh0 = (oh + 1) & 1;
                    const int p_a = (oh + 1 - kh0) >> 1, p_b = p_a - 1;        // patch rows reaching this pixel (kh0, kh0 + 2)
                    float a0 = bz0, a1 = bz1, a2 = bz2;
#pragma unroll
                    for (int dh = 0; dh < 2; ++dh) {
                        const int p = dh ? p_b : p_a;
                        if (dh ? (p_b < 0) : (p_a >= prm.P)) continue;
                        const float* srow = s_s + (((p % slots) << prm.lg_q)) * kDiSPitch + ((kh0 + 2 * dh) * 4 + kw0) * 3;
                        if (qa_ok) {
                            const float* src = srow + q_a * kDiSPitch;
                            a0 += src[0]; a1 += src[1]; a2 += src[2];
                        }
                        if (qb_ok) {
                            const float* src = srow + q_b * kDiSPitch + 6;      // kw0 + 2
                            a0 += src[0]; a1 += src[1]; a2 += src[2];
                        }
                    }
                    float* o = obase + (static_cast<long long>(oh) << lg_w) * 3;
                    o[0] = a0; o[1] = a1; o[2] = a2;
                    if (fuse) {
                        float* u = s_u + (((oh & (urows - 1)) << lg_w) + ow) * 3;
                        u[0] = a0; u[1] = a1; u[2] = a2;
                    }
                }
                if (fuse) {
                    named_bar(3, kDiEpi);
                    // 3x3 conv + tanh on the image rows whose three input rows are complete
                    const int c_lo = (p0 == 0) ? 0 : oh_lo - 1;
                    const int c_hi = last ? H - 1 : oh_hi - 1;
                    float* ibase = prm.img + (static_cast<long long>(n) * H * W + ow) * 3;
                    for (int oh = c_lo + row_first; oh <= c_hi; oh += row_step) {
                        float a0 = s_w[81], a1 = s_w[82], a2 = s_w[83];
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const int ih = oh + kh - 1;
                            if (ih < 0 || ih >= H) continue;
                            const float* urow = s_u + ((ih & (urows - 1)) << lg_w) * 3;
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                const int iw = ow + kw - 1;
                                if (iw < 0 || iw >= W) continue;
                                const float* u = urow + iw * 3;
                                const float* wp = s_w + (kh * 3 + kw) * 9;
#pragma unroll
                                for (int ci = 0; ci < 3; ++ci) {
                                    const float xv = u[ci];
                                    a0 += xv * wp[ci * 3 + 0];
                                    a1 += xv * wp[ci * 3 + 1];
                                    a2 += xv * wp[ci * 3 + 2];
                                }
                            }
                        }
                        float* o = ibase + (static_cast<long long>(oh) << lg_w) * 3;
                        o[0] = tanhf(a0); o[1] = tanhf(a1); o[2] = tanhf(a2);
                    }
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// ------------------------------------------------------------------------------------------------------------
// y[r][o] = act(sum_k x[r][k] * w[o][k] + b[o]), everything fp32.  CTA tile 16 rows x 32 outputs.  The operands of a
// tile (16 + 32 rows of up to 1024 floats = 192 KB) are fetched into shared memory AT ONCE with one bulk copy per row and
// K quarter -- a single exposure to the memory latency instead of one per K chunk (the 2 MB of operands are cold in a
// training step: the weights were last touched by Adam) -- and each of the four 64-thread groups starts on its K quarter
// as soon as that quarter has landed; the partial sums meet in shared memory.  A thread owns rows 2*tr, 2*tr + 1 and
// outputs to, to + 8, to + 16, to + 24 (row pitch K + 4 floats: 16-byte loads along K without bank conflicts).
constexpr int kDfR = 16, kDfO = 32, kDfG = 4, kDfKMax = 1024;
__global__ void __launch_bounds__(64 * kDfG) dense_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* y, int rows, int cin,
                                                              int cout, int act) {
    extern __shared__ __align__(16) uint8_t df_smem[];
    pdl_launch_dependents();
    pdl_wait();
    const int pitch = cin + 4;                               // floats per staged row
    float* xs = reinterpret_cast<float*>(df_smem);           // [16][pitch]
    float* ws = xs + kDfR * pitch;                           // [32][pitch]
    float* red = ws + kDfO * pitch;                          // [3][64][9]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + (kDfG - 1) * 64 * 9 + 1);   // 8-byte aligned: see the host's byte count
    bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~uintptr_t(7));
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int tr = t >> 3, to = t & 7;
    const int r0 = blockIdx.y * kDfR, o0 = blockIdx.x * kDfO;
    const int kq = ((cin + 4 * kDfG - 1) / (4 * kDfG)) * 4;  // K per group, a multiple of 4
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDfG; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < kDfG) {        // one issuing thread per K quarter
        const int q = threadIdx.x;
        const int k0 = q * kq, kn = min(cin, k0 + kq) - k0;
        int n_rows = 0;
        for (int r = 0; r < kDfR + kDfO; ++r) n_rows += (r < kDfR) ? (r0 + r < rows) : (o0 + r - kDfR < cout);
        mbar_arrive_expect_tx(&bars[q], kn > 0 ? static_cast<uint32_t>(n_rows * kn * 4) : 0u);
        if (kn > 0) {
            for (int r = 0; r < kDfR; ++r)
                if (r0 + r < rows) bulk_load_1d(xs + r * pitch + k0, x + (long long)(r0 + r) * cin + k0, kn * 4, &bars[q]);
            for (int o = 0; o < kDfO; ++o)
                if (o0 + o < cout) bulk_load_1d(ws + o * pitch + k0, w + (long long)(o0 + o) * cin + k0, kn * 4, &bars[q]);
        }
    }
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    {
        const int k0 = g * kq, k1 = min(cin, k0 + kq);
        mbar_wait(&bars[g], 0, 900 + g);
        const float* x0 = xs + (2 * tr) * pitch;
        const float* x1 = x0 + pitch;
        const float* wr[4] = {ws + to * pitch, ws + (to + 8) * pitch, ws + (to + 16) * pitch, ws + (to + 24) * pitch};
#pragma unroll 4
        for (int k = k0; k < k1; k += 4) {
            const float4 a = *reinterpret_cast<const float4*>(x0 + k);
            const float4 b = *reinterpret_cast<const float4*>(x1 + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 c = *reinterpret_cast<const float4*>(wr[j] + k);
                acc[0][j] += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
                acc[1][j] += b.x * c.x + b.y * c.y + b.z * c.z + b.w * c.w;
            }
        }
    }
    if (g > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[((g - 1) * 64 + t) * 9 + i] = acc[i >> 2][i & 3];
    }
    __syncthreads();
    if (g == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = r0 + 2 * tr + i;
            if (r >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = o0 + to + 8 * j;
                if (o < cout) {
                    float v = acc[i][j] + (bias != nullptr ? bias[o] : 0.f);
#pragma unroll
                    for (int gg = 0; gg < kDfG - 1; ++gg) v += red[(gg * 64 + t) * 9 + i * 4 + j];
                    if (act == T2I_ACT_LRELU) v = fmaxf(v, 0.2f * v);
                    else if (act == T2I_ACT_RELU) v = fmaxf(v, 0.f);
                    y[(long long)r * cout + o] = v;
                }
            }
        }
    }
}

}  // namespace t2i

using namespace t2i;

extern "C" int t2i_deconv_img(const t2i_act* a, const void* w, long long w_plane_stride, int w_rows, int w_cols, int w_layout,
                              int np, const float* bias3, float* out, const float* w9, const float* b9, float* img,
                              void* stream_) {
    if (a == nullptr || a->ptr == nullptr || w == nullptr || out == nullptr) return fail(T2I_ERR_BAD_ARG, "null tensor");
    if (np != 1 && np != 2) return fail(T2I_ERR_BAD_ARG, "np must be 1 or 2");
    if ((w9 != nullptr) != (img != nullptr) || (w9 != nullptr) != (b9 != nullptr))
        return fail(T2I_ERR_BAD_ARG, "w9, b9 and img go together");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DeconvImgParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.N = a->n; prm.P = a->h; prm.Q = a->w;
    if (prm.Q < 1 || (prm.Q & (prm.Q - 1)) != 0 || prm.Q > 128 || (prm.P * prm.Q) % 128 != 0)
        return fail(T2I_ERR_BAD_ARG, "deconv_img needs a power-of-two width <= 128 and h*w a multiple of 128 (got %d x %d)", prm.P,
                    prm.Q);
    prm.bp = 128 / prm.Q;
    if (prm.P % prm.bp != 0) return fail(T2I_ERR_BAD_ARG, "deconv_img: %d rows are not a multiple of %d", prm.P, prm.bp);
    prm.tiles_per_img = prm.P / prm.bp;
    for (prm.lg_q = 0; (1 << prm.lg_q) < prm.Q; ++prm.lg_q) {}
    for (prm.urows = 8; prm.urows < 2 * prm.bp + 3; prm.urows *= 2) {}
    prm.slots = prm.bp + 1;
    const bool kn = w_layout == T2I_W_KN;
    const int w_n = kn ? w_cols : w_rows, w_k = kn ? w_rows : w_cols;
    if (w_n != 64 || a->c > w_k || a->c > 256 || a->c % 8 != 0 || a->pitch % 8 != 0 || a->coff % 8 != 0 || w_cols % 8 != 0)
        return fail(T2I_ERR_BAD_ARG, "deconv_img: weights must be 64 (48 used) x K with K >= a.c (a.c=%d, weights %d x %d)", a->c,
                    w_rows, w_cols);
    prm.k_chunks = ceil_div(a->c, 64);
    prm.last_k_steps = ceil_div(a->c - (prm.k_chunks - 1) * 64, 16);
    prm.np = np;
    prm.n_pass = np == 2 ? 3 : 1;
    prm.b_kn = kn ? 1 : 0;
    prm.bias3 = bias3; prm.out = out; prm.w9 = w9; prm.b9 = b9; prm.img = img;
    const uint64_t e = 2;
    {
        const uint64_t plane_bytes = (np == 2) ? (uint64_t)a->plane_stride * e : (uint64_t)a->n * a->h * a->w * a->pitch * e;
        const uint64_t dims[5] = {(uint64_t)a->c, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n, (uint64_t)np};
        const uint64_t str[4] = {(uint64_t)a->pitch * e, (uint64_t)a->w * a->pitch * e, (uint64_t)a->h * a->w * a->pitch * e,
                                 plane_bytes};
        const uint32_t box[5] = {64, (uint32_t)prm.Q, (uint32_t)prm.bp, 1, 1};
        int rc = encode_tmap_bf16(&prm.a_map, static_cast<const __nv_bfloat16*>(a->ptr) + a->coff, 5, dims, str, box);
        if (rc != T2I_OK) return rc;
    }
    {
        const uint64_t plane_bytes = (np == 2) ? (uint64_t)w_plane_stride * e : (uint64_t)w_rows * w_cols * e;
        const uint64_t dims[3] = {(uint64_t)w_cols, (uint64_t)w_rows, (uint64_t)np};
        const uint64_t str[2] = {(uint64_t)w_cols * e, plane_bytes};
        const uint32_t box[3] = {64, 64, 1};
        int rc = encode_tmap_bf16(&prm.b_map, w, 3, dims, str, box);
        if (rc != T2I_OK) return rc;
    }
    const bool fuse = w9 != nullptr;
    const int smem_bytes = kDiStages * kDiABytes + np * prm.k_chunks * kDiBChunk + prm.slots * prm.Q * kDiSPitch * 4 +
                           (fuse ? prm.urows * 2 * prm.Q * 3 * 4 : 0) + 96 * 4 + 256 + 1024;
    if (smem_bytes > 232448) return fail(T2I_ERR_BAD_ARG, "deconv_img: shared memory plan of %d bytes", smem_bytes);
    static int attr_bytes = 0;      // the opt-in size is kept at what the plan needs: the occupancy the driver grants follows it
    if (attr_bytes != smem_bytes) {
        cudaError_t ce = cudaFuncSetAttribute(deconv_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (ce != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
        // two CTAs of ~110 KB per SM need the whole unified L1 / shared array as shared memory
        cudaFuncSetAttribute(deconv_img_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_bytes = smem_bytes;
    }
    // The gather / 3x3 epilogue runs one phase at a time on eight warps (IPC < 1): at the step's sizes the plan is ~110 KB
    // and TWO CTAs share an SM, each sweeping its own images -- one CTA's epilogue overlaps the other's loads and MMAs.
    // Decided from the shared-memory plan (320 threads x 84 registers and 128 TMEM columns per CTA fit twice); the
    // occupancy query answers 1 for this kernel, yet two CTAs per SM measured 64 -> 47 us (g.up4 forward) and 35 -> 27 us
    // (d.h0 input-gradient) against one per SM, i.e. they do run side by side.
    int per_sm = 1;
    {
        int dev = 0, sm_bytes = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_bytes, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        if (2 * (smem_bytes + 1024) <= sm_bytes) per_sm = 2;
    }
    const int slots_total = per_sm * num_sms();
    const int grid = prm.N < slots_total ? prm.N : slots_total;
    cudaError_t le = launch_pdl(deconv_img_kernel, grid, kDiThreads, smem_bytes, stream, prm);
    if (le != cudaSuccess) return fail(T2I_ERR_CUDA, "deconv_img_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("deconv_img_kernel");
}

extern "C" int t2i_dense_f32(const float* x, int rows, int cin, const float* w, const float* bias, int cout, int act, float* y,
                             void* stream) {
    if (x == nullptr || w == nullptr || y == nullptr) return fail(T2I_ERR_BAD_ARG, "null tensor");
    if (cin % 4 != 0 || cin > kDfKMax) return fail(T2I_ERR_BAD_ARG, "dense_f32: cin=%d must be a multiple of 4, at most %d", cin, kDfKMax);
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(w) & 15) != 0)
        return fail(T2I_ERR_BAD_ARG, "dense_f32: operands must be 16-byte aligned");
    const int smem_bytes = (kDfR + kDfO) * (cin + 4) * 4 + (kDfG - 1) * 64 * 9 * 4 + 16 + kDfG * 8;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t ce = cudaFuncSetAttribute(dense_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (kDfR + kDfO) * (kDfKMax + 4) * 4 + (kDfG - 1) * 64 * 9 * 4 + 16 + kDfG * 8);
        if (ce != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
        attr_done = true;
    }
    dim3 grid(ceil_div(cout, kDfO), ceil_div(rows, kDfR));
    cudaError_t le = launch_ew(dense_f32_kernel, grid, dim3(64 * kDfG), smem_bytes, static_cast<cudaStream_t>(stream), x, w, bias, y,
                               rows, cin, cout, act);
    if (le != cudaSuccess) return fail(T2I_ERR_CUDA, "dense_f32_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("dense_f32_kernel");
}
