// Implicit-GEMM convolution for sm_100a: TMA box loads (tap-shifted, OOB zero fill = padding)
// -> 128B-swizzled shared memory -> tcgen05.mma (M = 128 pixels, N = 128/256 channels, K = 64
// channels per step) -> fp32 accumulators in TMEM (double buffered) -> fused epilogue
// (bias, residual add, LeakyReLU/ReLU, derivative mask) -> bf16 planes in HBM.
//
// One persistent CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (each owns the 32 TMEM lanes its warp id % 4 selects).
//
// Replaces the TF ops behind utils/ops.py:61,69,87 of the reference (Conv2D,
// Conv2DBackpropInput, MatMul + BiasAdd + activation) for every dense contraction on the
// wgancls path, forward and input-gradient (see include/t2i_b200.h: t2i_conv_gemm).
#include "host_util.h"
#include "ptx.cuh"

namespace t2i {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB

struct EpiTensor {
    const __nv_bfloat16* ptr;
    long long plane_stride;
    int pitch, coff;
};

struct alignas(64) ConvGemmParams {
    CUtensorMap a_maps[4];
    CUtensorMap b_map;
    TapTable tt;
    int N, P, Q;           // virtual pixel grid (one GEMM row per point)
    int bn, bp, bq;        // TMA box on that grid, bn*bp*bq == 128
    int tiles_p, tiles_q;  // tiles along p and q (tiles along n = tiles_m / (tiles_p*tiles_q))
    int tiles_m, tiles_co, total_tiles;
    int k_chunks;          // ceil(Cin / 64)
    int n_pass;            // 1 (bf16) or 3 (split bf16: hi*hi, lo*hi, hi*lo)
    int np;
    int Cout;
    int OH, OW, osp, osq;  // output pixel = (p*osp + op, q*osq + oq)
    __nv_bfloat16* out;
    long long out_plane_stride;
    int out_pitch, out_coff;
    const float* bias;
    EpiTensor add, mask;
    int act, mask_kind;
};

template <int BLOCK_N>
struct SmemLayout {
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kBytes = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

// B_KN = false: weights [tap][N][K], K contiguous (K-major B operand, forward layout used forward).
// B_KN = true : weights [tap][K][N], N contiguous (MN-major B operand): the SAME packed forward
//               weights serve the input-gradient convolutions, no transposed copy exists.
template <int BLOCK_N, bool B_KN>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams prm) {
    using L = SmemLayout<BLOCK_N>;
    constexpr int kStages = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < prm.tt.n_maps; ++i) tma_prefetch_desc(&prm.a_maps[i]);
        tma_prefetch_desc(&prm.b_map);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * BLOCK_N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tpp = prm.tt.taps_per_phase;
    const int kb_per_tile = prm.n_pass * tpp * prm.k_chunks;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (one elected lane)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
                const int ct = tile % prm.tiles_co;
                const int rest = tile / prm.tiles_co;
                const int mt = rest % prm.tiles_m;
                const int ph = rest / prm.tiles_m;
                const int tq = mt % prm.tiles_q;
                const int tp = (mt / prm.tiles_q) % prm.tiles_p;
                const int tn = mt / (prm.tiles_q * prm.tiles_p);
                const int q0 = tq * prm.bq, p0 = tp * prm.bp, n0 = tn * prm.bn;
                for (int pass = 0; pass < prm.n_pass; ++pass) {
                    const int pa = (pass == 1) ? 1 : 0;  // A plane: hi, lo, hi
                    const int pb = (pass == 2) ? 1 : 0;  // B plane: hi, hi, lo
                    for (int t = 0; t < tpp; ++t) {
                        const Tap tap = prm.tt.taps[ph * tpp + t];
                        for (int kc = 0; kc < prm.k_chunks; ++kc) {
                            mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                            uint8_t* sa = smem + stage * L::kStageBytes;
                            tma_load_5d(&prm.a_maps[tap.map], &full_bar[stage], sa, kc * kBlockK, q0 + tap.dq,
                                        p0 + tap.dp, n0, pa);
                            if (!B_KN) {
                                tma_load_4d(&prm.b_map, &full_bar[stage], sa + kABytes, kc * kBlockK, ct * BLOCK_N,
                                            tap.wtap, pb);
                            } else {
#pragma unroll
                                for (int a = 0; a < BLOCK_N / 64; ++a)   // 64-channel atoms of [64 K rows x 128 B]
                                    tma_load_4d(&prm.b_map, &full_bar[stage], sa + kABytes + a * 8192,
                                                ct * BLOCK_N + a * 64, kc * kBlockK, tap.wtap, pb);
                            }
                            if (++stage == kStages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one elected lane)
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, B_KN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 300 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < kb_per_tile; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 200 + stage);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint64_t da = make_sw128_desc(sa, 0, 1024);
                    // K-major: +32 bytes per 16-element K step inside the 128B swizzle row (address >> 4);
                    // MN-major: 16 K rows of 128 B per step, 8 KB between 64-channel atoms.
                    const uint64_t db = B_KN ? make_sw128_desc(sa + kABytes, 8192, 1024) : make_sw128_desc(sa + kABytes, 0, 1024);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        umma_bf16(d_tmem, da + 2 * k, db + (B_KN ? 128 * k : 2 * k), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == kb_per_tile - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------ epilogue warps
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int np = prm.np;
        int it = 0;
        for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int ct = tile % prm.tiles_co;
            const int rest = tile / prm.tiles_co;
            const int mt = rest % prm.tiles_m;
            const int ph = rest / prm.tiles_m;
            const int tq = mt % prm.tiles_q;
            const int tp = (mt / prm.tiles_q) % prm.tiles_p;
            const int tn = mt / (prm.tiles_q * prm.tiles_p);
            const int q = tq * prm.bq + row % prm.bq;
            const int p = tp * prm.bp + (row / prm.bq) % prm.bp;
            const int n = tn * prm.bn + row / (prm.bq * prm.bp);
            const bool valid = (n < prm.N) && (p < prm.P) && (q < prm.Q);
            const long long pix = (static_cast<long long>(n) * prm.OH + (p * prm.osp + prm.tt.ph_op[ph])) * prm.OW +
                                  (q * prm.osq + prm.tt.ph_oq[ph]);
            const int co_base = ct * BLOCK_N;

            mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                if (co_base + c0 >= prm.Cout) break;  // warp-uniform
                __syncwarp();
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                if (!valid) continue;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int co = co_base + c0 + g * 8;
                    if (co >= prm.Cout) break;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                    if (prm.bias != nullptr) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm.bias + co));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(prm.bias + co + 4));
                        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                    }
                    if (prm.add.ptr != nullptr) {
                        const __nv_bfloat16* ap = prm.add.ptr + pix * prm.add.pitch + prm.add.coff + co;
                        for (int pl = 0; pl < np; ++pl) {
                            const uint4 u = *reinterpret_cast<const uint4*>(ap + pl * prm.add.plane_stride);
                            v[0] += bf16_lo(u.x); v[1] += bf16_hi(u.x); v[2] += bf16_lo(u.y); v[3] += bf16_hi(u.y);
                            v[4] += bf16_lo(u.z); v[5] += bf16_hi(u.z); v[6] += bf16_lo(u.w); v[7] += bf16_hi(u.w);
                        }
                    }
                    if (prm.act == T2I_ACT_LRELU) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);
                    } else if (prm.act == T2I_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                    }
                    if (prm.mask.ptr != nullptr) {
                        const __nv_bfloat16* mp = prm.mask.ptr + pix * prm.mask.pitch + prm.mask.coff + co;
                        float m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        for (int pl = 0; pl < np; ++pl) {
                            const uint4 u = *reinterpret_cast<const uint4*>(mp + pl * prm.mask.plane_stride);
                            m[0] += bf16_lo(u.x); m[1] += bf16_hi(u.x); m[2] += bf16_lo(u.y); m[3] += bf16_hi(u.y);
                            m[4] += bf16_lo(u.z); m[5] += bf16_hi(u.z); m[6] += bf16_lo(u.w); m[7] += bf16_hi(u.w);
                        }
                        const float neg = (prm.mask_kind == T2I_MASK_LRELU) ? 0.2f : 0.0f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] *= (m[j] > 0.0f) ? 1.0f : neg;
                    }
                    __nv_bfloat16* op = prm.out + pix * prm.out_pitch + prm.out_coff + co;
                    uint4 hi;
                    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
                    hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
                    *reinterpret_cast<uint4*>(op) = hi;
                    if (np == 2) {
                        uint4 lo;
                        lo.x = pack_bf16x2(v[0] - bf16_lo(hi.x), v[1] - bf16_hi(hi.x));
                        lo.y = pack_bf16x2(v[2] - bf16_lo(hi.y), v[3] - bf16_hi(hi.y));
                        lo.z = pack_bf16x2(v[4] - bf16_lo(hi.z), v[5] - bf16_hi(hi.z));
                        lo.w = pack_bf16x2(v[6] - bf16_lo(hi.w), v[7] - bf16_hi(hi.w));
                        *reinterpret_cast<uint4*>(op + prm.out_plane_stride) = lo;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BLOCK_N);
    }
}

static int make_act_maps(const t2i_act& x, int mode, int np, int bq, int bp, int bn, CUtensorMap* maps, int n_maps) {
    if (x.pitch % 8 != 0 || x.coff % 8 != 0 || x.c % 8 != 0)
        return fail(T2I_ERR_BAD_ARG, "activation channels must be multiples of 8 (c=%d pitch=%d coff=%d)", x.c,
                    x.pitch, x.coff);
    const uint64_t e = 2;  // bytes per element
    const uint64_t plane_bytes = (np == 2) ? (uint64_t)x.plane_stride * e : (uint64_t)x.n * x.h * x.w * x.pitch * e;
    const uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)bq, (uint32_t)bp, (uint32_t)bn, 1};
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(x.ptr) + x.coff;
    if (n_maps == 1) {
        const uint64_t dims[5] = {(uint64_t)x.c, (uint64_t)x.w, (uint64_t)x.h, (uint64_t)x.n, (uint64_t)np};
        const uint64_t str[4] = {(uint64_t)x.pitch * e, (uint64_t)x.w * x.pitch * e, (uint64_t)x.h * x.w * x.pitch * e,
                                 plane_bytes};
        return encode_tmap_bf16(&maps[0], base, 5, dims, str, box);
    }
    // stride-2 parity views: view (rh, rw) holds pixels (2*p + rh, 2*q + rw)
    if ((x.h & 1) || (x.w & 1)) return fail(T2I_ERR_BAD_ARG, "K4S2 needs even h, w (got %d x %d)", x.h, x.w);
    for (int rh = 0; rh < 2; ++rh)
        for (int rw = 0; rw < 2; ++rw) {
            const uint64_t dims[5] = {(uint64_t)x.c, (uint64_t)x.w / 2, (uint64_t)x.h / 2, (uint64_t)x.n, (uint64_t)np};
            const uint64_t str[4] = {2 * (uint64_t)x.pitch * e, 2 * (uint64_t)x.w * x.pitch * e,
                                     (uint64_t)x.h * x.w * x.pitch * e, plane_bytes};
            int rc = encode_tmap_bf16(&maps[rh * 2 + rw], base + ((long long)rh * x.w + rw) * x.pitch, 5, dims, str, box);
            if (rc != T2I_OK) return rc;
        }
    return T2I_OK;
}

// Box on the virtual grid with bq*bp*bn == rows (a power of two).
static void choose_box(int N, int P, int Q, int rows, int* bn, int* bp, int* bq) {
    int q = floor_pow2(Q);
    if (q > rows) q = rows;
    int p = floor_pow2(P);
    if (p > rows / q) p = rows / q;
    *bq = q;
    *bp = p;
    *bn = rows / (q * p);
    (void)N;
}

static EpiTensor epi_of(const t2i_act& a) {
    EpiTensor e;
    e.ptr = static_cast<const __nv_bfloat16*>(a.ptr);
    e.plane_stride = a.plane_stride;
    e.pitch = a.pitch;
    e.coff = a.coff;
    return e;
}

}  // namespace t2i

using namespace t2i;

extern "C" int t2i_conv_gemm(const t2i_conv_gemm_desc* d, void* stream_) {
    if (d == nullptr) return fail(T2I_ERR_BAD_ARG, "null descriptor");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    int rc = build_taps(d->mode, d->k, d->flip, &prm.tt);
    if (rc != T2I_OK) return rc;
    if (d->np != 1 && d->np != 2) return fail(T2I_ERR_BAD_ARG, "np must be 1 or 2");
    const t2i_act& x = d->x;
    const t2i_act& y = d->y;
    if (x.ptr == nullptr || y.ptr == nullptr || d->w == nullptr) return fail(T2I_ERR_BAD_ARG, "null tensor");
    // weight matrix per tap: w_rows x w_cols, cols contiguous.  NK: rows = output channels, cols = contraction;
    // KN: rows = contraction, cols = output channels.
    const bool kn = d->w_layout == T2I_W_KN;
    const int w_n = kn ? d->w_cols : d->w_rows, w_k = kn ? d->w_rows : d->w_cols;
    if (x.c > w_k || d->w_cols % 8 != 0) return fail(T2I_ERR_BAD_ARG, "x.c=%d vs weight contraction=%d (cols %d)", x.c, w_k, d->w_cols);
    if (y.c > w_n || y.c % 8 != 0 || y.pitch % 8 != 0 || y.coff % 8 != 0)
        return fail(T2I_ERR_BAD_ARG, "bad output channels c=%d pitch=%d coff=%d weight n=%d", y.c, y.pitch, y.coff, w_n);
    // virtual grid and output mapping
    prm.N = x.n;
    if (d->mode == T2I_CONV_S1) {
        prm.P = x.h; prm.Q = x.w; prm.OH = x.h; prm.OW = x.w; prm.osp = prm.osq = 1;
    } else if (d->mode == T2I_CONV_K4S2) {
        prm.P = x.h / 2; prm.Q = x.w / 2; prm.OH = x.h / 2; prm.OW = x.w / 2; prm.osp = prm.osq = 1;
    } else {
        prm.P = x.h; prm.Q = x.w; prm.OH = 2 * x.h; prm.OW = 2 * x.w; prm.osp = prm.osq = 2;
    }
    if (y.n != x.n || y.h != prm.OH || y.w != prm.OW)
        return fail(T2I_ERR_BAD_ARG, "output shape [%d,%d,%d] does not match expected [%d,%d,%d]", y.n, y.h, y.w, x.n,
                    prm.OH, prm.OW);
    choose_box(prm.N, prm.P, prm.Q, kBlockM, &prm.bn, &prm.bp, &prm.bq);
    prm.tiles_q = ceil_div(prm.Q, prm.bq);
    prm.tiles_p = ceil_div(prm.P, prm.bp);
    prm.tiles_m = ceil_div(prm.N, prm.bn) * prm.tiles_p * prm.tiles_q;
    prm.Cout = y.c;
    prm.k_chunks = ceil_div(x.c, kBlockK);
    prm.np = d->np;
    prm.n_pass = (d->np == 2) ? 3 : 1;
    const int sms = num_sms();
    int block_n = 128;
    if (y.c > 128 && (long long)prm.tt.n_phases * prm.tiles_m * ceil_div(y.c, 256) >= sms) block_n = 256;
    prm.tiles_co = ceil_div(y.c, block_n);
    prm.total_tiles = prm.tt.n_phases * prm.tiles_m * prm.tiles_co;

    rc = make_act_maps(x, d->mode, d->np, prm.bq, prm.bp, prm.bn, prm.a_maps, prm.tt.n_maps);
    if (rc != T2I_OK) return rc;
    {
        const int taps = (d->mode == T2I_CONV_S1) ? d->k * d->k : 16;
        const uint64_t e = 2;
        const uint64_t plane_bytes = (d->np == 2) ? (uint64_t)d->w_plane_stride * e : (uint64_t)taps * d->w_rows * d->w_cols * e;
        const uint64_t dims[4] = {(uint64_t)d->w_cols, (uint64_t)d->w_rows, (uint64_t)taps, (uint64_t)d->np};
        const uint64_t str[3] = {(uint64_t)d->w_cols * e, (uint64_t)d->w_cols * d->w_rows * e, plane_bytes};
        const uint32_t box[4] = {64, kn ? 64u : (uint32_t)block_n, 1, 1};
        rc = encode_tmap_bf16(&prm.b_map, d->w, 4, dims, str, box);
        if (rc != T2I_OK) return rc;
    }
    prm.out = static_cast<__nv_bfloat16*>(y.ptr);
    prm.out_plane_stride = y.plane_stride;
    prm.out_pitch = y.pitch;
    prm.out_coff = y.coff;
    prm.bias = d->bias;
    prm.add = epi_of(d->add);
    prm.mask = epi_of(d->mask);
    prm.act = d->act;
    prm.mask_kind = d->mask_kind;
    if (d->add.ptr && (d->add.pitch % 8 || d->add.coff % 8)) return fail(T2I_ERR_BAD_ARG, "add tensor misaligned");
    if (d->mask.ptr && (d->mask.pitch % 8 || d->mask.coff % 8)) return fail(T2I_ERR_BAD_ARG, "mask tensor misaligned");
    if (d->mask.ptr && d->mask_kind == T2I_MASK_NONE) return fail(T2I_ERR_BAD_ARG, "mask tensor without mask_kind");

    const int grid = prm.total_tiles < sms ? prm.total_tiles : sms;
    typedef void (*KernelFn)(const ConvGemmParams);
    static bool attr_done[4] = {false, false, false, false};
    const int variant = (block_n == 256 ? 2 : 0) + (kn ? 1 : 0);
    const KernelFn fns[4] = {conv_gemm_kernel<128, false>, conv_gemm_kernel<128, true>, conv_gemm_kernel<256, false>,
                             conv_gemm_kernel<256, true>};
    const int smem_bytes = block_n == 256 ? SmemLayout<256>::kBytes : SmemLayout<128>::kBytes;
    if (!attr_done[variant]) {
        cudaError_t e = cudaFuncSetAttribute(fns[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_done[variant] = true;
    }
    fns[variant]<<<grid, kThreads, smem_bytes, stream>>>(prm);
    return check_launch("conv_gemm_kernel");
}
