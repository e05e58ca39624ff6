// Implicit-GEMM convolution for sm_100a: TMA box loads (tap-shifted, OOB zero fill = padding)
// -> 128B-swizzled shared memory -> tcgen05.mma (M = 128 pixels, N = 128/256 channels, K = 64
// channels per step) -> fp32 accumulators in TMEM (double buffered) -> fused epilogue
// (bias, residual add, LeakyReLU/ReLU, derivative mask) -> bf16 planes staged in swizzled shared
// memory -> TMA tensor stores.  The residual / mask tiles the epilogue needs are themselves
// TMA-loaded (prefetched one or two 64-channel sub-tiles ahead), so every global access of the
// kernel is a full-line bulk transfer.
//
// CTA2 = true pairs two CTAs of a cluster (cta_group::2): a 256-pixel x N tile, each CTA stages its own
// 128 pixel rows and half of the weight tile, the leader CTA issues M = 256 MMAs, each CTA runs the
// epilogue of its own 128 accumulator rows.  Operand bytes per MMA cycle drop by a third.
//
// One persistent CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..9 = epilogue (a warp reaches the 32 TMEM lanes its id % 4 selects; two warps share each
// lane quarter and split the 64 columns of a sub-tile between them).
//
// Replaces the TF ops behind utils/ops.py:61,69,87 of the reference (Conv2D,
// Conv2DBackpropInput, MatMul + BiasAdd + activation) for every dense contraction on the
// wgancls path, forward and input-gradient (see include/t2i_b200.h: t2i_conv_gemm).
#include "host_util.h"
#include "img_patch.cuh"
#include "ptx.cuh"

namespace t2i {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 320;                  // 2 control warps + 8 epilogue warps
// A_IMG: two or four more warps assemble the patch rows of the A tile (four where the register file allows it:
// the STATS epilogue needs 168 registers per thread, 448 threads of that do not fit)
__host__ __device__ constexpr int img_producers(bool stats) { return stats ? 64 : 128; }
constexpr int kImgRing = 4;                    // A_IMG: raw image-row buffers in flight (cp.async ring, tiles ahead)
constexpr int kEpiThreads = 256;
constexpr int kEpiGroup = 128;                 // an epilogue group: four warps, thread <-> accumulator row
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kSubBytes = kBlockM * 64 * 2;     // one 128-pixel x 64-channel bf16 sub-tile = 16 KB
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 232448 - 1024;      // 227 KB minus alignment slack

struct alignas(64) ConvGemmParams {
    CUtensorMap a_maps[4];     // input views (stride-2 convs: four parity views)
    CUtensorMap b_map;         // packed weights
    CUtensorMap out_maps[4];   // output views (transposed convs: one parity view per phase)
    CUtensorMap add_maps[4];   // optional residual, same grid as the output
    CUtensorMap mask_maps[4];  // optional derivative-mask source, same grid as the output
    CUtensorMap sx_maps[4];    // optional second factor of the per-channel dot statistic, same grid as the output
    TapTable tt;
    int N, P, Q;               // virtual pixel grid (one GEMM row per point)
    int bn, bp, bq;            // TMA box on that grid, bn*bp*bq == 128
    int tiles_p, tiles_q;
    int tiles_m, tiles_co, total_tiles;
    FastDiv fd_co, fd_ph, fd_q, fd_p;   // divisors of the tile index: tiles_co, n_phases, tiles_q, tiles_p
    int lg_bq, lg_bqp;         // log2(bq), log2(bq * bp): row -> (q, p, n) inside the (power-of-two) box
    int k_chunks;              // ceil(Cin / 64)
    int n_pass;                // 1 (bf16) or 3 (split bf16: hi*hi, lo*hi, hi*lo)
    int np;
    int Cout;
    int n_stages;              // mainloop pipeline depth (what fits beside the epilogue buffers)
    int epi_depth;             // 1, 2 or 4 staging buffers per epilogue tensor and group
    int epi_groups;            // 2: two epilogue groups; 1: group 1 idles (np = 2: shared memory)
    int epi_split;             // 1: the groups share every tile (sub-tiles g, g + 2, ..); 0: they take alternate tiles
    int has_add, has_mask, has_sx;
    const float* bias;
    int act, mask_kind;
    // per-channel statistics of the output values (fp32 atomics): sum, and sum of squares (second_kind 1)
    // or dot with the sx tile (second_kind 2); rows of samples >= stat_n / channels >= stat_c are left out
    float* stat_sum;
    float* stat_second;
    int second_kind;
    int stat_n, stat_c;
    int w_n0;                  // first output channel inside the weight matrix (output-channel window)
    int last_k_steps;          // 16-channel MMA steps of the LAST K chunk (input channels beyond x.c are zero fill: skipped)
    int stat_acc;              // channels of per-CTA shared accumulators (flushed once at the end), 0 = none
    int dbg;                   // T2I_STAT_DBG bit field (tools/bench_conv.py): skip parts of the statistics path
    // A_IMG: the A operand is the 4x4 / stride-2 patch matrix of this fp32 NHWC 3-channel image (row = output pixel,
    // column (kh*4 + kw)*3 + c, 48 columns), assembled in shared memory by two producer warps -- never in HBM
    const uint32_t* img;       // padded bf16 rows [np][N][img_h][pitch] as 32-bit words (t2i_img_to_rows)
    long long img_plane_words; // words between the planes
    int img_h, img_w;
    // debug build + T2I_TIMELINE=1 (tools/conv_timeline.py): CTA 0 records %globaltimer at the pipeline's hand-over points of its
    // first 64 tiles: [tile][event], events 0 weights TMA issued, 1 patch rows arrived, 2 MMA sees the stage, 3 MMA
    // committed, 4 epilogue sees the accumulator, 5 epilogue done with the tile, 6 MMA got a free accumulator
    unsigned long long* timeline;
};
#ifdef T2I_TIMELINE_BUILD     // make -C csrc EXTRA=-DT2I_TIMELINE_BUILD: a DEBUG build -- the stamps in the MMA issue loop cost 10-25 %
#define T2I_MARK(it, ev) \
    do { if (prm.timeline != nullptr && blockIdx.x == 0 && (it) < 64) prm.timeline[(it) * 8 + (ev)] = global_timer_ns(); } while (0)
#else
#define T2I_MARK(it, ev) do {} while (0)
#endif

// Column totals of a 32 x 32 block held one row per lane, 32 values per lane: after the five folds lane l
// holds the total of column l (31 shuffles instead of 32 x 5).
template <int STEP>
__device__ __forceinline__ void fold_cols(float (&s)[32], int lane) {
    const bool upper = (lane & STEP) != 0;
#pragma unroll
    for (int i = 0; i < STEP; ++i) {
        const float send = upper ? s[i] : s[i + STEP];
        const float keep = upper ? s[i + STEP] : s[i];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, STEP);
    }
}
__device__ __forceinline__ float warp_col_totals(float (&s)[32], int lane) {
    fold_cols<16>(s, lane);
    fold_cols<8>(s, lane);
    fold_cols<4>(s, lane);
    fold_cols<2>(s, lane);
    fold_cols<1>(s, lane);
    return s[0];
}

__device__ __forceinline__ void tma_store_5d(const void* tmap, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct TileCoord {
    int ct, ph, q0, p0, n0;
};
// `tile` indexes (pixel tile [or pixel-tile pair], phase, channel tile); pair > 1 selects this CTA's
// pixel tile inside the pair by its cluster rank.
__device__ __forceinline__ TileCoord decode_tile(const ConvGemmParams& prm, int tile, int pair = 1, int rank = 0) {
    TileCoord t;
    // order: output-channel tile fastest, then phase, then pixel tile -- the four phases of a transposed
    // conv re-read the same input rows, so they run back to back and hit in L2
    const int rest = fast_div(tile, prm.fd_co);
    t.ct = tile - rest * prm.tiles_co;
    const int rest2 = fast_div(rest, prm.fd_ph);
    t.ph = rest - rest2 * prm.tt.n_phases;
    const int mt = rest2 * pair + rank;
    const int mq = fast_div(mt, prm.fd_q);
    t.q0 = (mt - mq * prm.tiles_q) * prm.bq;
    const int mp = fast_div(mq, prm.fd_p);
    t.p0 = (mq - mp * prm.tiles_p) * prm.bp;
    t.n0 = mp * prm.bn;
    return t;
}

// B_KN = false: weights [tap][N][K], K contiguous (K-major B operand, forward layout used forward).
// B_KN = true : weights [tap][K][N], N contiguous (MN-major B operand): the SAME packed forward
//               weights serve the input-gradient convolutions, no transposed copy exists.
// STATS = true adds the per-channel statistics of the epilogue (a separate instantiation so that plain
// launches keep the lean epilogue).
// A_IMG = true (single CTA, BLOCK_N = 128): the 3-channel ends.  One K block of 48 (= 3 MMA steps) per tile and pass;
// warp 0 fetches only the weights, warps 10-11 (10-13 without STATS) stage the 2*bp + 2 image rows of the tile in shared memory (coalesced
// 16-byte loads, zero rows = SAME padding) and write each pixel's 48 patch values as bf16 into the 128B-swizzled
// K-major A tile (fence.proxy.async, then they arrive on the stage's full barrier next to the weight TMA).
template <int BLOCK_N, bool B_KN, bool CTA2, bool STATS, bool A_IMG = false>
__global__ void __launch_bounds__(A_IMG ? kThreads + img_producers(STATS) : kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams prm) {
    constexpr int kImgProducers = img_producers(STATS);
    static_assert(!A_IMG || (!CTA2 && BLOCK_N == 128), "the image-patch producer exists for single-CTA 128-wide tiles");
    constexpr int kBRows = CTA2 ? BLOCK_N / 2 : BLOCK_N;     // weight rows (output channels) staged by this CTA
    constexpr int kBBytes = kBRows * kBlockK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int kPair = CTA2 ? 2 : 1;
    // accumulator buffers in TMEM (512 columns): with four, the MMAs run up to three tiles ahead of the epilogue, which
    // hides the commit -> wait -> load -> release handshakes of short-K tiles (they cost more than the MMAs themselves)
    constexpr int kAccBufs = (BLOCK_N <= 128) ? 4 : 2;
    const int rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;
    const int unit0 = blockIdx.x / kPair;                     // first tile (pair) of this CTA (pair)
    const int unit_stride = gridDim.x / kPair;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int n_stages = prm.n_stages;
    const int D = prm.epi_depth;
    const int np = prm.np;
    // carve-up: [stages][out G*D*np][add G*D*np][mask G*D*np][sx G*D*np][bias G][stat partials G][totals G][barriers]
    // (G = epilogue groups, each with its own staging buffers)
    const int G = prm.epi_groups;
    uint8_t* s_out = smem + n_stages * kStageBytes;
    uint8_t* s_add = s_out + G * D * np * kSubBytes;
    uint8_t* s_mask = s_add + (prm.has_add ? G * D * np * kSubBytes : 0);
    uint8_t* s_sx = s_mask + (prm.has_mask ? G * D * np * kSubBytes : 0);
    float* s_bias = reinterpret_cast<float*>(s_sx + (prm.has_sx ? G * D * np * kSubBytes : 0));
    float* s_stat = s_bias + 2 * BLOCK_N;           // STATS: per group [2 statistics][2 column halves][4 lane quarters][32]
    float* s_acc = s_stat + (STATS ? 1024 : 0);     // STATS: per group running totals [2][stat_acc] (0 = straight to global)
    // A_IMG: ring of kImgRing slots, each np x [2*bp + 2] padded bf16 image rows (bulk copies, tiles ahead)
    uint32_t* s_img = reinterpret_cast<uint32_t*>(s_acc + (STATS ? 4 * prm.stat_acc : 0));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_img + (A_IMG ? kImgRing * np * (2 * prm.bp + 2) * img_pitch_words(prm.img_w) : 0));
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 4;
    uint64_t* aux_full = tmem_empty + 4;
    uint64_t* rows_full = aux_full + 4;          // A_IMG: one per ring slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rows_full + kImgRing);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();   // the next kernel may run its prologue under this grid's tail
    if (threadIdx.x == 0) T2I_MARK(63, 0);      // row 63 of the timeline: kernel-level stamps

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < prm.tt.n_maps; ++i) tma_prefetch_desc(&prm.a_maps[i]);
        tma_prefetch_desc(&prm.b_map);
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(&full_bar[i], A_IMG ? 1 + kImgProducers : kPair);   // one arrival per producer (pair: in the leader)
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kAccBufs; ++i) {
            mbar_init(&tmem_full[i], 1);
            // the warps that read the tile (one group, or both when they share tiles), in both CTAs (leader's copy)
            mbar_init(&tmem_empty[i], kPair * (kEpiGroup / 32) * (prm.epi_split ? prm.epi_groups : 1));
        }
        for (int i = 0; i < 4; ++i) mbar_init(&aux_full[i], 1);
        for (int i = 0; i < kImgRing; ++i) mbar_init(&rows_full[i], 1);
        fence_barrier_init();
    }
    if (CTA2) cluster_sync_all();               // peer barriers exist before anything can target them
    if (warp == 1) {
        if (CTA2) {
            tmem_alloc_2sm(tmem_slot, kAccBufs * BLOCK_N);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, kAccBufs * BLOCK_N);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) T2I_MARK(63, 1);
    pdl_wait();                // prologue done; from here on global memory of earlier kernels is read
    if (threadIdx.x == 0) T2I_MARK(63, 2);

    const int tpp = prm.tt.taps_per_phase;
    const int kb_per_tile = prm.n_pass * tpp * prm.k_chunks;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (one elected lane)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = unit0; tile < prm.total_tiles; tile += unit_stride) {
                const TileCoord tc = decode_tile(prm, tile, kPair, rank);
                for (int pass = 0; pass < prm.n_pass; ++pass) {
                    const int pa = (pass == 1) ? 1 : 0;  // A plane: hi, lo, hi
                    const int pb = (pass == 2) ? 1 : 0;  // B plane: hi, hi, lo
                    for (int t = 0; t < tpp; ++t) {
                        const Tap tap = prm.tt.taps[tc.ph * tpp + t];
                        for (int kc = 0; kc < prm.k_chunks; ++kc) {
                            mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                            uint8_t* sa = smem + stage * kStageBytes;
                            const int brow = prm.w_n0 + tc.ct * BLOCK_N + rank * kBRows;   // this CTA's slice of the weight tile
                            if (A_IMG) {       // the weights only; the A tile comes from the producer warps
                                mbar_arrive_expect_tx(&full_bar[stage], kBBytes);
                                T2I_MARK((tile - unit0) / unit_stride, 0);
                                if (!B_KN) {
                                    tma_load_4d(&prm.b_map, &full_bar[stage], sa + kABytes, kc * kBlockK, brow, tap.wtap, pb);
                                } else {
#pragma unroll
                                    for (int a = 0; a < kBRows / 64; ++a)
                                        tma_load_4d(&prm.b_map, &full_bar[stage], sa + kABytes + a * 8192, brow + a * 64,
                                                    kc * kBlockK, tap.wtap, pb);
                                }
                            } else if (!CTA2) {
                                mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                                tma_load_5d(&prm.a_maps[tap.map], &full_bar[stage], sa, kc * kBlockK, tc.q0 + tap.dq,
                                            tc.p0 + tap.dp, tc.n0, pa);
                                if (!B_KN) {
                                    tma_load_4d(&prm.b_map, &full_bar[stage], sa + kABytes, kc * kBlockK, brow, tap.wtap, pb);
                                } else {
#pragma unroll
                                    for (int a = 0; a < kBRows / 64; ++a)   // 64-channel atoms of [64 K rows x 128 B]
                                        tma_load_4d(&prm.b_map, &full_bar[stage], sa + kABytes + a * 8192, brow + a * 64,
                                                    kc * kBlockK, tap.wtap, pb);
                                }
                            } else {
                                // both producers arrive on the LEADER's barrier; it expects the bytes of both CTAs
                                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
                                else mbar_arrive_cluster(&full_bar[stage], 0);
                                tma_load_5d_2sm(&prm.a_maps[tap.map], &full_bar[stage], sa, kc * kBlockK, tc.q0 + tap.dq,
                                                tc.p0 + tap.dp, tc.n0, pa);
                                if (!B_KN) {
                                    tma_load_4d_2sm(&prm.b_map, &full_bar[stage], sa + kABytes, kc * kBlockK, brow, tap.wtap, pb);
                                } else {
#pragma unroll
                                    for (int a = 0; a < kBRows / 64; ++a)
                                        tma_load_4d_2sm(&prm.b_map, &full_bar[stage], sa + kABytes + a * 8192, brow + a * 64,
                                                        kc * kBlockK, tap.wtap, pb);
                                }
                            }
                            if (++stage == n_stages) {
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
        if (rank == 0 && elect_one()) {         // in a CTA pair only the leader issues
            constexpr uint32_t idesc = make_idesc_bf16(kPair * kBlockM, BLOCK_N, 0, B_KN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = unit0; tile < prm.total_tiles; tile += unit_stride, ++it) {
                const int acc = it & (kAccBufs - 1);
                const uint32_t acc_phase = (it / kAccBufs) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 300 + acc);
                tc_fence_after();
                T2I_MARK(it, 6);
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                int kc = 0;     // K chunk inside the tap: the last one may hold fewer than 64 real input channels
                for (int kb = 0; kb < kb_per_tile; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 200 + stage);
                    tc_fence_after();
                    if (kb == 0) T2I_MARK(it, 2);
                    const int k_steps = (kc == prm.k_chunks - 1) ? prm.last_k_steps : kBlockK / 16;
                    if (++kc == prm.k_chunks) kc = 0;
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = make_sw128_desc(sa, 0, 1024);
                    // K-major: +32 bytes per 16-element K step inside the 128B swizzle row (address >> 4);
                    // MN-major: 16 K rows of 128 B per step, 8 KB between 64-channel atoms.
                    const uint64_t db = B_KN ? make_sw128_desc(sa + kABytes, 8192, 1024) : make_sw128_desc(sa + kABytes, 0, 1024);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        if (k < k_steps) {
                            if (CTA2) umma_bf16_2sm(d_tmem, da + 2 * k, db + (B_KN ? 128 * k : 2 * k), idesc, (kb | k) != 0);
                            else umma_bf16(d_tmem, da + 2 * k, db + (B_KN ? 128 * k : 2 * k), idesc, (kb | k) != 0);
                        }
                    }
                    if (CTA2) {     // free the stage / publish the accumulator in BOTH CTAs
                        umma_commit_2sm(&empty_bar[stage]);
                        if (kb == kb_per_tile - 1) umma_commit_2sm(&tmem_full[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == kb_per_tile - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (kb == kb_per_tile - 1) T2I_MARK(it, 3);
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (A_IMG && warp >= 10) {
        // ------------------------------------------------ image-patch producers (64 threads, two tile rows each)
        const int pt = threadIdx.x - kThreads;
        const int n_rows = 2 * prm.bp + 2;
        const int pitchw = img_pitch_words(prm.img_w);
        const int slot_words = np * n_rows * pitchw;
        const long long sample_words = static_cast<long long>(prm.img_h) * pitchw;
        // the rows of the next kImgRing - 1 tiles are in flight (bulk copies) while this tile's patches are assembled:
        // one tile's work is far shorter than the memory latency
        auto fetch = [&](int tile, int slot) {
            if (tile >= prm.total_tiles) return;
            const TileCoord tn = decode_tile(prm, tile, 1, 0);
            img_rows_fetch(prm.img, prm.img_plane_words, sample_words, np, tn.n0, 2 * tn.p0 - 1, prm.img_h, n_rows, pitchw,
                           s_img + slot * slot_words, &rows_full[slot], pt, kImgProducers);
        };
        int stage = 0, slot = 0, round = 0;
        uint32_t phase = 0;
        for (int a = 0; a < kImgRing - 1; ++a) fetch(unit0 + a * unit_stride, a);
        for (int tile = unit0; tile < prm.total_tiles; tile += unit_stride) {
            const TileCoord tc = decode_tile(prm, tile, 1, 0);
            named_bar_sync(5, kImgProducers);     // everybody is done with the previous tile: its ring slot is free
            fetch(tile + (kImgRing - 1) * unit_stride, (slot + kImgRing - 1) & (kImgRing - 1));
            mbar_wait(&rows_full[slot], round & 1, 800 + slot);
            const uint32_t* s_bf = s_img + slot * slot_words;
            const uint32_t* s_lo = s_bf + n_rows * pitchw;
            for (int pass = 0; pass < prm.n_pass; ++pass) {
                const uint32_t* plane = (pass == 1) ? s_lo : s_bf;             // A plane: hi, lo, hi
                mbar_wait(&empty_bar[stage], phase ^ 1, 700 + stage);
                uint8_t* sa = smem + stage * kStageBytes;
#pragma unroll 1
                for (int half_r = 0; half_r < kBlockM / kImgProducers; ++half_r) {
                    const int row = pt + half_r * kImgProducers;
                    int q = tc.q0 + (row & (prm.bq - 1));                      // bn == 1: rows = (pl, q)
                    if (q >= prm.Q) q = prm.Q - 1;                             // ragged tile: the row is clipped by the store
                    img_patch_row(plane, pitchw, row >> prm.lg_bq, q, sa + (row >> 3) * 1024 + (row & 7) * 128, row, false);
                }
                fence_proxy_async();                           // generic-proxy writes -> visible to the tensor core
                mbar_arrive(&full_bar[stage]);
                if (pt == 0 && pass == 0) T2I_MARK((tile - unit0) / unit_stride, 1);
                if (++stage == n_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++slot == kImgRing) {
                slot = 0;
                ++round;
            }
        }
    } else if (!A_IMG || warp < 10) {
        // ------------------------------------------------ epilogue: two groups of four warps (thread <-> accumulator row)
        // Group g takes this CTA's tiles g, g + 2, ...: while one group waits for a barrier, a TMA store or an accumulator,
        // the other one works, and every named barrier spans 128 threads.  A thread walks the 64 columns of a sub-tile
        // as two halves of 32.
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int et = threadIdx.x - 64 - grp * kEpiGroup;     // thread index inside the group
        const int row = quarter * 32 + lane;
        const bool leader = (((warp - 2) & 3) == 0 && lane == 0);
        const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;     // the group's named barriers
        if (grp < G) {
        s_out += grp * D * np * kSubBytes; s_add += grp * D * np * kSubBytes;
        s_mask += grp * D * np * kSubBytes; s_sx += grp * D * np * kSubBytes;
        s_bias += grp * BLOCK_N; s_stat += grp * 512; s_acc += grp * 2 * prm.stat_acc;
        aux_full += grp * 2;
        const bool has_aux = prm.has_add || prm.has_mask || prm.has_sx;
        const uint32_t aux_bytes = static_cast<uint32_t>((prm.has_add + prm.has_mask + prm.has_sx) * np * kSubBytes);
        if (STATS) {   // running per-group totals: thread et owns channels == et mod 64 of statistic et / 64
            for (int i = et; i < 2 * prm.stat_acc; i += kEpiGroup) s_acc[i] = 0.f;
        }
        const int sw = row & 7;                       // 128B swizzle: 16B chunk j of row r lives at chunk j ^ (r & 7)
        const uint32_t row_off = row * 128;
        const float neg = (prm.mask_kind == T2I_MASK_LRELU) ? 0.2f : 0.0f;
        // shared addresses of this thread's row in the staging tiles, and its four 16-byte chunks inside the row
        const uint32_t so_u32 = smem_u32(s_out) + row_off, sa_u32 = smem_u32(s_add) + row_off;
        const uint32_t sm_u32 = smem_u32(s_mask) + row_off, sx_u32 = smem_u32(s_sx) + row_off;
        const uint32_t sb0_u32 = smem_u32(s_bias);

        // aux tiles are prefetched D sub-tiles ahead along the sequence (tile, sub) this CTA will process
        auto issue_aux = [&](int tile, int sub, int buf) {
            const TileCoord tc = decode_tile(prm, tile, kPair, rank);
            const int c0 = tc.ct * BLOCK_N + sub * 64;
            mbar_arrive_expect_tx(&aux_full[buf], aux_bytes);
            for (int pl = 0; pl < np; ++pl) {
                if (prm.has_add)
                    tma_load_5d(&prm.add_maps[tc.ph], &aux_full[buf], s_add + (buf * np + pl) * kSubBytes, c0, tc.q0, tc.p0,
                                tc.n0, pl);
                if (prm.has_mask)
                    tma_load_5d(&prm.mask_maps[tc.ph], &aux_full[buf], s_mask + (buf * np + pl) * kSubBytes, c0, tc.q0,
                                tc.p0, tc.n0, pl);
                if (prm.has_sx)
                    tma_load_5d(&prm.sx_maps[tc.ph], &aux_full[buf], s_sx + (buf * np + pl) * kSubBytes, c0, tc.q0,
                                tc.p0, tc.n0, pl);
            }
        };
        auto subs_of = [&](int tile) {
            const int ct = tile - fast_div(tile, prm.fd_co) * prm.tiles_co;
            int rem = prm.Cout - ct * BLOCK_N;
            if (rem > BLOCK_N) rem = BLOCK_N;
            return (rem + 63) / 64;
        };
        // The group's sequence of (tile, sub-tile): alternate tiles with all their sub-tiles, or (epi_split) every tile
        // with sub-tiles grp, grp + G, ...  The aux prefetch cursor runs D elements ahead of the compute cursor.
        const bool split = prm.epi_split != 0;
        const int tile_step = split ? unit_stride : G * unit_stride, it_step = split ? 1 : G;
        const int tile_first = unit0 + (split ? 0 : grp * unit_stride), it_first = split ? 0 : grp;
        const int sub_first = split ? grp : 0, sub_step = split ? G : 1;
        int pf_tile = tile_first, pf_sub = sub_first;
        while (pf_tile < prm.total_tiles && pf_sub >= subs_of(pf_tile)) pf_tile += tile_step;
        auto pf_advance = [&]() {
            pf_sub += sub_step;
            if (pf_sub >= subs_of(pf_tile)) {
                pf_sub = sub_first;
                pf_tile += tile_step;
                while (pf_tile < prm.total_tiles && pf_sub >= subs_of(pf_tile)) pf_tile += tile_step;
            }
        };
        if (has_aux && leader) {
            for (int i = 0; i < D; ++i) {
                if (pf_tile < prm.total_tiles) {
                    issue_aux(pf_tile, pf_sub, i);
                    pf_advance();
                }
            }
        }

        int g = 0;   // running index of 64-channel sub-tiles processed by this group
        int it = it_first;
        for (int tile = tile_first; tile < prm.total_tiles; tile += tile_step, it += it_step) {
            const int acc = it & (kAccBufs - 1);
            const uint32_t acc_phase = (it / kAccBufs) & 1;
            const TileCoord tc = decode_tile(prm, tile, kPair, rank);
            const int co_base = tc.ct * BLOCK_N;
            const int n_sub = subs_of(tile);
            const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
            // statistics: does this tile contribute at all, and is this thread's row a real, counted pixel?
            const bool tile_stats = STATS && tc.n0 < prm.stat_n && co_base < prm.stat_c;
            // ragged tiles (box beyond the grid, or straddling the sample limit) mask their rows individually
            const bool tile_ragged = STATS && (tc.n0 + prm.bn > prm.stat_n || tc.p0 + prm.bp > prm.P || tc.q0 + prm.bq > prm.Q);
            bool row_ok = true;
            if (tile_ragged)
                row_ok = tc.n0 + (row >> prm.lg_bqp) < prm.stat_n && tc.p0 + ((row >> prm.lg_bq) & (prm.bp - 1)) < prm.P &&
                         tc.q0 + (row & (prm.bq - 1)) < prm.Q;
            if (leader) T2I_MARK(it, 7);
            // stage this tile's bias slice (all threads passed the previous sub-tile's second barrier)
            if (prm.bias != nullptr) {
                for (int i = et; i < BLOCK_N; i += kEpiGroup) s_bias[i] = (co_base + i < prm.Cout) ? __ldg(prm.bias + co_base + i) : 0.f;
            }
            if (sub_first >= n_sub) {       // shared tile without a sub-tile for this group: stay in step, release it
                mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
                __syncwarp();
                if (lane == 0) {
                    if (CTA2 && rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
                    else mbar_arrive(&tmem_empty[acc]);
                }
                continue;
            }
            for (int sub = sub_first; sub < n_sub; sub += sub_step, ++g) {
                const int buf = g & (D - 1);                       // D is 1, 2 or 4
                const uint32_t aux_parity = (D == 4) ? ((g >> 2) & 1) : (D == 2) ? ((g >> 1) & 1) : (g & 1);
                if (leader) {   // the store that last used out[buf] must have finished reading it
                    if (D == 4) bulk_wait_read<3>(); else if (D == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                }
                if (has_aux) mbar_wait(&aux_full[buf], aux_parity, 500 + buf);
                named_bar_sync(bar_a, kEpiGroup);
                if (sub == sub_first) {
                    mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
                    tc_fence_after();
                    if (leader) T2I_MARK(it, 4);
                }
                uint8_t* o_hi = s_out + (buf * np) * kSubBytes + row_off;
                const uint8_t* a_base = s_add + (buf * np) * kSubBytes + row_off;
                const uint8_t* m_base = s_mask + (buf * np) * kSubBytes + row_off;
                const uint8_t* x_base = s_sx + (buf * np) * kSubBytes + row_off;
                // the two 32-column halves of the sub-tile; without the statistics' extra registers both halves are
                // in flight together (two independent dependency chains per warp)
#pragma unroll(STATS ? 1 : 2)
                for (int half = 0; half < 2; ++half) {
                    const uint32_t sb_u32 = sb0_u32 + static_cast<uint32_t>(half * 32) * 4;
                    uint32_t coff4[4];
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) coff4[gq] = static_cast<uint32_t>(((half * 4 + gq) ^ sw) * 16);
                    __syncwarp();
                    uint32_t r[32];
                    float r2[32];     // second statistic's per-element terms (only live when requested)
                    tmem_ld_32x32(taddr + sub * 64 + half * 32, r);
                    tmem_ld_wait();
                    if (np == 1) {
                        // Throughput path (single bf16 plane): every optional stage is ONE uniform branch around
                        // straight-line code over the thread's 32 values (independent chains keep the two epilogue
                        // warps of a scheduler issuing back to back), shared memory through 32-bit addresses.
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        const uint32_t boff = static_cast<uint32_t>(buf) * kSubBytes;
                        if (prm.bias != nullptr) {
                            const uint32_t sb = sb_u32 + static_cast<uint32_t>(sub) * 256;
#pragma unroll
                            for (int q4 = 0; q4 < 8; ++q4) {
                                const uint4 b = lds128(sb + q4 * 16);
                                v[4 * q4 + 0] += __uint_as_float(b.x); v[4 * q4 + 1] += __uint_as_float(b.y);
                                v[4 * q4 + 2] += __uint_as_float(b.z); v[4 * q4 + 3] += __uint_as_float(b.w);
                            }
                        }
                        if (prm.has_add) {
#pragma unroll
                            for (int gq = 0; gq < 4; ++gq) {
                                const uint4 u = lds128(sa_u32 + boff + coff4[gq]);
                                v[gq * 8 + 0] += bf16_lo(u.x); v[gq * 8 + 1] += bf16_hi(u.x);
                                v[gq * 8 + 2] += bf16_lo(u.y); v[gq * 8 + 3] += bf16_hi(u.y);
                                v[gq * 8 + 4] += bf16_lo(u.z); v[gq * 8 + 5] += bf16_hi(u.z);
                                v[gq * 8 + 6] += bf16_lo(u.w); v[gq * 8 + 7] += bf16_hi(u.w);
                            }
                        }
                        if (prm.act == T2I_ACT_LRELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.2f * v[i]);
                        } else if (prm.act == T2I_ACT_RELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
                        }
                        if (prm.has_mask) {
                            // bf16 m > 0  <=>  its bits, as a signed integer, are > 0 (low half shifted up; high half
                            // compared against 0xffff so that the low half does not matter)
#pragma unroll
                            for (int gq = 0; gq < 4; ++gq) {
                                const uint4 u = lds128(sm_u32 + boff + coff4[gq]);
                                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float a = v[gq * 8 + 2 * k], b = v[gq * 8 + 2 * k + 1];
                                    v[gq * 8 + 2 * k] = (static_cast<int>(w4[k] << 16) > 0) ? a : a * neg;
                                    v[gq * 8 + 2 * k + 1] = (static_cast<int>(w4[k]) > 0xffff) ? b : b * neg;
                                }
                            }
                        }
                        if (tile_stats) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
                            if (prm.second_kind == 1) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) r2[i] = v[i] * v[i];
                            } else if (prm.second_kind == 2) {
#pragma unroll
                                for (int gq = 0; gq < 4; ++gq) {
                                    const uint4 u = lds128(sx_u32 + boff + coff4[gq]);
                                    r2[gq * 8 + 0] = v[gq * 8 + 0] * bf16_lo(u.x); r2[gq * 8 + 1] = v[gq * 8 + 1] * bf16_hi(u.x);
                                    r2[gq * 8 + 2] = v[gq * 8 + 2] * bf16_lo(u.y); r2[gq * 8 + 3] = v[gq * 8 + 3] * bf16_hi(u.y);
                                    r2[gq * 8 + 4] = v[gq * 8 + 4] * bf16_lo(u.z); r2[gq * 8 + 5] = v[gq * 8 + 5] * bf16_hi(u.z);
                                    r2[gq * 8 + 6] = v[gq * 8 + 6] * bf16_lo(u.w); r2[gq * 8 + 7] = v[gq * 8 + 7] * bf16_hi(u.w);
                                }
                            }
                        }
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            uint4 hi;
                            hi.x = pack_bf16x2(v[gq * 8 + 0], v[gq * 8 + 1]); hi.y = pack_bf16x2(v[gq * 8 + 2], v[gq * 8 + 3]);
                            hi.z = pack_bf16x2(v[gq * 8 + 4], v[gq * 8 + 5]); hi.w = pack_bf16x2(v[gq * 8 + 6], v[gq * 8 + 7]);
                            sts128(so_u32 + boff + coff4[gq], hi);
                        }
                    } else
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {   // parity path (np == 2: hi + lo planes), generic form
                        const int chunk = half * 4 + gq;              // 16B chunk (8 channels) within the 64-channel row
                        const uint32_t coff = static_cast<uint32_t>((chunk ^ sw) * 16);
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                        if (prm.bias != nullptr) {
                            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + sub * 64 + chunk * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + sub * 64 + chunk * 8 + 4);
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                        }
                        if (prm.has_add) {
                            for (int pl = 0; pl < np; ++pl) {
                                const uint4 u = *reinterpret_cast<const uint4*>(a_base + pl * kSubBytes + coff);
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
                        if (prm.has_mask) {
                            float m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                            for (int pl = 0; pl < np; ++pl) {
                                const uint4 u = *reinterpret_cast<const uint4*>(m_base + pl * kSubBytes + coff);
                                m[0] += bf16_lo(u.x); m[1] += bf16_hi(u.x); m[2] += bf16_lo(u.y); m[3] += bf16_hi(u.y);
                                m[4] += bf16_lo(u.z); m[5] += bf16_hi(u.z); m[6] += bf16_lo(u.w); m[7] += bf16_hi(u.w);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= (m[j] > 0.0f) ? 1.0f : neg;
                        }
                        if (tile_stats) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) r[gq * 8 + j] = __float_as_uint(v[j]);
                            if (prm.second_kind == 1) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) r2[gq * 8 + j] = v[j] * v[j];
                            } else if (prm.second_kind == 2) {
                                float xs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                                for (int pl = 0; pl < np; ++pl) {
                                    const uint4 u = *reinterpret_cast<const uint4*>(x_base + pl * kSubBytes + coff);
                                    xs[0] += bf16_lo(u.x); xs[1] += bf16_hi(u.x); xs[2] += bf16_lo(u.y); xs[3] += bf16_hi(u.y);
                                    xs[4] += bf16_lo(u.z); xs[5] += bf16_hi(u.z); xs[6] += bf16_lo(u.w); xs[7] += bf16_hi(u.w);
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) r2[gq * 8 + j] = v[j] * xs[j];
                            }
                        }
                        uint4 hi;
                        hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
                        hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
                        *reinterpret_cast<uint4*>(o_hi + coff) = hi;
                        if (np == 2) {
                            uint4 lo;
                            lo.x = pack_bf16x2(v[0] - bf16_lo(hi.x), v[1] - bf16_hi(hi.x));
                            lo.y = pack_bf16x2(v[2] - bf16_lo(hi.y), v[3] - bf16_hi(hi.y));
                            lo.z = pack_bf16x2(v[4] - bf16_lo(hi.z), v[5] - bf16_hi(hi.z));
                            lo.w = pack_bf16x2(v[6] - bf16_lo(hi.w), v[7] - bf16_hi(hi.w));
                            *reinterpret_cast<uint4*>(o_hi + kSubBytes + coff) = lo;
                        }
                    }
                    if (tile_stats) {   // column totals of this warp's 32 x 32 block -> shared partials
                        float s1[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) s1[i] = __uint_as_float(r[i]);
                        if (!row_ok) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) { s1[i] = 0.f; r2[i] = 0.f; }
                        }
                        if (!(prm.dbg & 1)) {
                        s_stat[(half * 4 + quarter) * 32 + lane] = warp_col_totals(s1, lane);
                        if (prm.second_kind != 0) s_stat[256 + (half * 4 + quarter) * 32 + lane] = warp_col_totals(r2, lane);
                        }
                    }
                }
                if (sub + sub_step >= n_sub) {   // accumulator fully read by this group: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CTA2 && rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
                        else mbar_arrive(&tmem_empty[acc]);
                    }
                }
                fence_proxy_async();      // generic-proxy smem writes -> visible to the TMA store
                if (!A_IMG && leader && sub == 0) T2I_MARK(it, 0);
                named_bar_sync(bar_b, kEpiGroup);
                if (!A_IMG && leader && sub == 0) T2I_MARK(it, 1);
                if (tile_stats && !(prm.dbg & 2)) {   // (statistic, channel) totals of the sub-tile: 128 threads = 2 x 64
                    const int which = et >> 6, hc = et & 63;          // hc = column inside the 64-channel sub-tile
                    float* dst = which == 0 ? prm.stat_sum : prm.stat_second;
                    const int ch = co_base + sub * 64 + hc;
                    if (dst != nullptr && ch < prm.stat_c && (which == 0 || prm.second_kind != 0)) {
                        const float* sp = s_stat + which * 256 + (hc >> 5) * 128 + (hc & 31);
                        const float t = sp[0] + sp[32] + sp[64] + sp[96];
                        if (prm.stat_acc > 0) s_acc[which * prm.stat_acc + ch] += t;   // this thread alone owns the slot
                        else atomicAdd(dst + ch, t);
                    }
                }
                if (leader) {
                    for (int pl = 0; pl < np; ++pl)
                        tma_store_5d(&prm.out_maps[tc.ph], s_out + (buf * np + pl) * kSubBytes, co_base + sub * 64, tc.q0,
                                     tc.p0, tc.n0, pl);
                    bulk_commit();
                    if (has_aux && pf_tile < prm.total_tiles) {   // aux[buf] has been consumed by every thread
                        issue_aux(pf_tile, pf_sub, buf);
                        pf_advance();
                    }
                    if (sub + sub_step >= n_sub) T2I_MARK(it, 5);
                }
            }
        }
        // the stores must have left shared memory before the CTA exits; their global writes complete with the grid
        if (leader) bulk_wait_read<0>();
        if (leader && grp == 0) T2I_MARK(63, 6);
        if (STATS && prm.stat_acc > 0 && !(prm.dbg & 4)) {   // one global atomic per channel this group contributed to
            const int which = et >> 6;
            float* dst = which == 0 ? prm.stat_sum : prm.stat_second;
            if (dst != nullptr) {
                for (int ch = et & 63; ch < prm.stat_c; ch += 64) {
                    const float t = s_acc[which * prm.stat_acc + ch];
                    if (t != 0.f) atomicAdd(dst + ch, t);
                }
            }
        }
        }   // grp < G
    }

    if (threadIdx.x == 0) T2I_MARK(63, 3);
    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();   // the peer may still be read by the leader's MMAs
    if (threadIdx.x == 0) T2I_MARK(63, 4);
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (CTA2) tmem_dealloc_2sm(tmem_base, kAccBufs * BLOCK_N);
        else tmem_dealloc(tmem_base, kAccBufs * BLOCK_N);
        if (lane == 0) T2I_MARK(63, 5);
    }
}

// Tensor maps of an NHWC planes view on the kernel's pixel grid: one map, or the four stride-2
// parity views (view (rh, rw) holds pixels (2p + rh, 2q + rw)).
static int make_act_maps(const t2i_act& x, bool parity, int np, int bq, int bp, int bn, CUtensorMap* maps) {
    if (x.pitch % 8 != 0 || x.coff % 8 != 0 || x.c % 8 != 0)
        return fail(T2I_ERR_BAD_ARG, "activation channels must be multiples of 8 (c=%d pitch=%d coff=%d)", x.c,
                    x.pitch, x.coff);
    const uint64_t e = 2;  // bytes per element
    const uint64_t plane_bytes = (np == 2) ? (uint64_t)x.plane_stride * e : (uint64_t)x.n * x.h * x.w * x.pitch * e;
    const uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)bq, (uint32_t)bp, (uint32_t)bn, 1};
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(x.ptr) + x.coff;
    if (!parity) {
        const uint64_t dims[5] = {(uint64_t)x.c, (uint64_t)x.w, (uint64_t)x.h, (uint64_t)x.n, (uint64_t)np};
        const uint64_t str[4] = {(uint64_t)x.pitch * e, (uint64_t)x.w * x.pitch * e, (uint64_t)x.h * x.w * x.pitch * e,
                                 plane_bytes};
        return encode_tmap_bf16(&maps[0], base, 5, dims, str, box);
    }
    if ((x.h & 1) || (x.w & 1)) return fail(T2I_ERR_BAD_ARG, "stride-2 views need even h, w (got %d x %d)", x.h, x.w);
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
static void choose_box(int P, int Q, int rows, int* bn, int* bp, int* bq) {
    int q = floor_pow2(Q);
    if (q > rows) q = rows;
    int p = floor_pow2(P);
    if (p > rows / q) p = rows / q;
    *bq = q;
    *bp = p;
    *bn = rows / (q * p);
}

}  // namespace t2i

using namespace t2i;

static unsigned long long* g_timeline = nullptr;     // debug build: kTimelineRegions x [64 tiles][8 events], one region per launch (round robin)
static const int kTimelineRegions = 32;
static int g_timeline_next = 0;
extern "C" int t2i_debug_timeline(unsigned long long* host_dst, int count) {
    if (g_timeline == nullptr || count > kTimelineRegions * 64 * 8) return fail(T2I_ERR_BAD_ARG, "no timeline recorded (T2I_TIMELINE=1)");
    cudaError_t e = cudaMemcpy(host_dst, g_timeline, sizeof(unsigned long long) * count, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? T2I_OK : fail(T2I_ERR_CUDA, "cudaMemcpy: %s", cudaGetErrorString(e));
}

extern "C" int t2i_conv_gemm(const t2i_conv_gemm_desc* d, void* stream_) {
    if (d == nullptr) return fail(T2I_ERR_BAD_ARG, "null descriptor");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    const bool a_img = d->x_img != nullptr;
    int rc = T2I_OK;
    if (a_img) {
        // the image form: ONE contraction block of 48 = (kh, kw, c) per output pixel; the weights are [1][rows][cols]
        if (d->mode != T2I_CONV_K4S2 || d->k != 4 || d->flip != 0)
            return fail(T2I_ERR_BAD_ARG, "x_img goes with mode T2I_CONV_K4S2, k = 4");
        prm.tt.n_phases = 1; prm.tt.taps_per_phase = 1; prm.tt.n_maps = 0;
    } else {
        rc = build_taps(d->mode, d->k, d->flip, &prm.tt);
        if (rc != T2I_OK) return rc;
    }
    if (d->np != 1 && d->np != 2) return fail(T2I_ERR_BAD_ARG, "np must be 1 or 2");
    const t2i_act& x = d->x;
    const t2i_act& y = d->y;
    if ((x.ptr == nullptr && !a_img) || y.ptr == nullptr || d->w == nullptr) return fail(T2I_ERR_BAD_ARG, "null tensor");
    if (a_img && (x.c != 48 || x.w % 4 != 0)) return fail(T2I_ERR_BAD_ARG, "x_img: x.c must be 48 and the width a multiple of 4");
    if (a_img && (reinterpret_cast<uintptr_t>(d->x_img) & 15) != 0) return fail(T2I_ERR_BAD_ARG, "x_img must be 16-byte aligned");
    // weight matrix per tap: w_rows x w_cols, cols contiguous.  NK: rows = output channels, cols = contraction;
    // KN: rows = contraction, cols = output channels.
    const bool kn = d->w_layout == T2I_W_KN;
    const int w_n = kn ? d->w_cols : d->w_rows, w_k = kn ? d->w_rows : d->w_cols;
    if (x.c > w_k || d->w_cols % 8 != 0) return fail(T2I_ERR_BAD_ARG, "x.c=%d vs weight contraction=%d (cols %d)", x.c, w_k, d->w_cols);
    if (d->w_n0 < 0 || d->w_n0 % 8 != 0) return fail(T2I_ERR_BAD_ARG, "w_n0 must be a non-negative multiple of 8");
    prm.w_n0 = d->w_n0;
    if (d->w_n0 + y.c > w_n || y.c % 8 != 0 || y.pitch % 8 != 0 || y.coff % 8 != 0)
        return fail(T2I_ERR_BAD_ARG, "bad output channels c=%d pitch=%d coff=%d weight n=%d", y.c, y.pitch, y.coff, w_n);
    // virtual grid and output extent
    int OH, OW;
    prm.N = x.n;
    if (d->mode == T2I_CONV_S1) {
        prm.P = x.h; prm.Q = x.w; OH = x.h; OW = x.w;
    } else if (d->mode == T2I_CONV_K4S2) {
        prm.P = x.h / 2; prm.Q = x.w / 2; OH = x.h / 2; OW = x.w / 2;
    } else {
        prm.P = x.h; prm.Q = x.w; OH = 2 * x.h; OW = 2 * x.w;
    }
    if (y.n != x.n || y.h != OH || y.w != OW)
        return fail(T2I_ERR_BAD_ARG, "output shape [%d,%d,%d] does not match expected [%d,%d,%d]", y.n, y.h, y.w, x.n, OH, OW);
    choose_box(prm.P, prm.Q, kBlockM, &prm.bn, &prm.bp, &prm.bq);
    prm.tiles_q = ceil_div(prm.Q, prm.bq);
    prm.tiles_p = ceil_div(prm.P, prm.bp);
    prm.tiles_m = ceil_div(prm.N, prm.bn) * prm.tiles_p * prm.tiles_q;
    prm.Cout = y.c;
    prm.k_chunks = ceil_div(x.c, kBlockK);
    prm.last_k_steps = ceil_div(x.c - (prm.k_chunks - 1) * kBlockK, 16);
    prm.np = d->np;
    prm.n_pass = (d->np == 2) ? 3 : 1;
    prm.has_add = d->add.ptr != nullptr;
    prm.has_mask = d->mask.ptr != nullptr;
    if (prm.has_mask && d->mask_kind == T2I_MASK_NONE) return fail(T2I_ERR_BAD_ARG, "mask tensor without mask_kind");
    prm.has_sx = d->stat_dot != nullptr;
    if (prm.has_sx && d->stat_x.ptr == nullptr) return fail(T2I_ERR_BAD_ARG, "stat_dot without stat_x");
    if (d->stat_sq != nullptr && d->stat_dot != nullptr) return fail(T2I_ERR_BAD_ARG, "stat_sq and stat_dot are exclusive");
    prm.stat_sum = d->stat_sum;
    prm.stat_second = d->stat_sq != nullptr ? d->stat_sq : d->stat_dot;
    prm.second_kind = d->stat_sq != nullptr ? 1 : d->stat_dot != nullptr ? 2 : 0;
    prm.stat_n = (d->stat_n > 0 && d->stat_n < x.n) ? d->stat_n : x.n;
    prm.stat_c = (d->stat_c > 0 && d->stat_c < y.c) ? d->stat_c : y.c;
    for (const t2i_act* a : {&d->add, &d->mask, &d->stat_x}) {
        const int need_c = (a == &d->stat_x) ? prm.stat_c : y.c;     // the dot factor only has to cover the counted channels
        if (a->ptr != nullptr && (a->n != y.n || a->h != y.h || a->w != y.w || a->c < need_c))
            return fail(T2I_ERR_BAD_ARG, "epilogue tensor [%d,%d,%d,%d] does not cover the output [%d,%d,%d,%d]", a->n, a->h,
                        a->w, a->c, y.n, y.h, y.w, need_c);
    }
    const int sms = num_sms();
    // CTA pairs (cta_group::2, 256-pixel tiles) whenever there are at least two pixel tiles
    const bool allow_cta2 = [] { const char* e = getenv("T2I_CONV_CTA2"); return !(e && e[0] == '0'); }();
    // at most 64 output channels (the 128x128 / 256x256 maps of StackGAN stage-II and PGGAN): a 64-wide channel tile
    // halves the MMA work of the 128-wide one.  With MN-major weights a CTA pair would stage half a 64-channel
    // atom per CTA, so those launches stay single-CTA.
    static const bool allow_n64 = [] { const char* e = getenv("T2I_CONV_N64"); return !(e && e[0] == '0'); }();
    const bool small_n = allow_n64 && y.c <= 64 && !a_img;
    const bool cta2 = allow_cta2 && prm.tiles_m >= 2 && !(small_n && kn) && !a_img;
    if (a_img && prm.bn != 1) return fail(T2I_ERR_BAD_ARG, "x_img needs at least 128 output pixels per image (got %d x %d)", prm.P, prm.Q);
    const int units_m = cta2 ? ceil_div(prm.tiles_m, 2) : prm.tiles_m;     // schedulable pixel tiles (pairs)
    const int workers = cta2 ? sms / 2 : sms;                              // CTAs or CTA pairs
    int block_n = small_n ? 64 : 128;
    if (y.c > 128 && (long long)prm.tt.n_phases * units_m * ceil_div(y.c, 256) >= workers && !a_img) block_n = 256;
    if (!a_img) {   // tools/bench_conv.py: force the channel tile (development aid)
        const char* e = getenv("T2I_CONV_BN");
        if (e && atoi(e) == 256 && y.c > 128) block_n = 256;
        if (e && atoi(e) == 128 && !small_n) block_n = 128;
    }
    prm.tiles_co = ceil_div(y.c, block_n);
    prm.total_tiles = prm.tt.n_phases * units_m * prm.tiles_co;
    prm.fd_co = make_fastdiv(prm.tiles_co);
    prm.fd_ph = make_fastdiv(prm.tt.n_phases);
    prm.fd_q = make_fastdiv(prm.tiles_q);
    prm.fd_p = make_fastdiv(prm.tiles_p);
    for (prm.lg_bq = 0; (1 << prm.lg_bq) < prm.bq; ++prm.lg_bq) {}
    for (prm.lg_bqp = 0; (1 << prm.lg_bqp) < prm.bq * prm.bp; ++prm.lg_bqp) {}
    // shared memory plan: epilogue staging first, the mainloop pipeline takes what is left
    prm.epi_depth = (d->np == 1) ? 2 : 1;
    const int b_rows = cta2 ? block_n / 2 : block_n;                       // weight rows staged per CTA
    const int stage_bytes = kABytes + b_rows * kBlockK * 2;
    const int n_aux = prm.has_add + prm.has_mask + prm.has_sx;
    if (n_aux == 3) prm.epi_depth = 1;
    // two epilogue groups (each with its own staging buffers) in the throughput mode; with aux tiles one buffer each
    prm.epi_groups = (d->np == 1) ? 2 : 1;
    {
        static const int force = [] { const char* e = getenv("T2I_EPI_GROUPS"); return e ? atoi(e) : 0; }();
        if (force == 1 && n_aux >= 2) prm.epi_groups = 1;
    }
    // 256-wide tiles have two TMEM accumulators: both groups must finish a tile together (sub-tile split), or the
    // MMAs of the tile after next wait for a whole single-group epilogue
    // launches of at most one tile per CTA would leave the second group idle: there too both groups share the tile
    prm.epi_split = (block_n == 256 || prm.total_tiles <= workers) ? 1 : 0;
    {
        static const int force = [] { const char* e = getenv("T2I_EPI_MODE"); return e ? atoi(e) : 0; }();
        if (force == 1) prm.epi_split = 0;
        if (force == 2) prm.epi_split = 1;
    }
    // one staging buffer per group = the two of a single group, so that the mainloop keeps its stages; tiles of one or two
    // K blocks are bound by the epilogue (the wait for the previous store to leave the buffer is on its critical path)
    // and need few stages: two buffers per group there
    const bool short_k = prm.n_pass * prm.tt.taps_per_phase * prm.k_chunks <= 2;
    if (prm.epi_groups == 2) prm.epi_depth = (short_k && n_aux <= 1) ? 2 : 1;
    {
        static const int force = [] { const char* e = getenv("T2I_EPI_DEPTH"); return e ? atoi(e) : 0; }();
        if (force == 3 && a_img) prm.epi_depth = 2;
        if ((force == 1 || force == 2) && n_aux <= 1) prm.epi_depth = force;
    }
    const int epi_bytes = prm.epi_groups * (1 + n_aux) * prm.epi_depth * d->np * kSubBytes;
    const bool stats = prm.stat_sum != nullptr || prm.stat_second != nullptr;
    // per-CTA running totals in shared memory (one flush per CTA instead of one atomic per tile and channel:
    // thousands of tiles hammering the same few L2 lines serialise); very wide outputs have few tiles -> direct
    prm.stat_acc = (stats && prm.stat_c <= 2048) ? (prm.stat_c + 63) / 64 * 64 : 0;
    {
        const char* e = getenv("T2I_STAT_DBG");
        prm.dbg = e ? atoi(e) : 0;
        if (prm.dbg & 8) prm.stat_acc = 0;
    }
    {
        static const bool want = [] { const char* e = getenv("T2I_TIMELINE"); return e && e[0] == '1'; }();
        if (want && g_timeline == nullptr) {
            cudaMalloc(&g_timeline, kTimelineRegions * 64 * 8 * sizeof(unsigned long long));
            cudaMemset(g_timeline, 0, kTimelineRegions * 64 * 8 * sizeof(unsigned long long));
        }
        prm.timeline = want ? g_timeline + (g_timeline_next++ % kTimelineRegions) * 64 * 8 : nullptr;
    }
    prm.img = static_cast<const uint32_t*>(d->x_img);
    prm.img_plane_words = x.plane_stride / 2;
    prm.img_h = x.h;
    prm.img_w = x.w;
    const int img_bytes = a_img ? kImgRing * d->np * (2 * prm.bp + 2) * img_pitch_words(x.w) * 4 : 0;   // the ring of padded bf16 rows
    const int tail_bytes = 2 * block_n * 4 + (stats ? 2 * (2048 + 8 * prm.stat_acc) : 0) + img_bytes + 512;   // bias, statistics (per group), image rows, barriers
    int stages = (kSmemBudget - epi_bytes - tail_bytes) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return fail(T2I_ERR_BAD_ARG, "shared memory plan leaves %d pipeline stages", stages);
    prm.n_stages = stages;
    const int smem_bytes = stages * stage_bytes + epi_bytes + tail_bytes + 1024;

    const bool parity_in = d->mode == T2I_CONV_K4S2, parity_out = d->mode == T2I_DECONV_K4S2;
    if (!a_img) {
        rc = make_act_maps(x, parity_in, d->np, prm.bq, prm.bp, prm.bn, prm.a_maps);
        if (rc != T2I_OK) return rc;
    }
    rc = make_act_maps(y, parity_out, d->np, prm.bq, prm.bp, prm.bn, prm.out_maps);
    if (rc != T2I_OK) return rc;
    if (prm.has_add) {
        t2i_act a = d->add;
        a.c = y.c;
        rc = make_act_maps(a, parity_out, d->np, prm.bq, prm.bp, prm.bn, prm.add_maps);
        if (rc != T2I_OK) return rc;
    }
    if (prm.has_mask) {
        t2i_act a = d->mask;
        a.c = y.c;
        rc = make_act_maps(a, parity_out, d->np, prm.bq, prm.bp, prm.bn, prm.mask_maps);
        if (rc != T2I_OK) return rc;
    }
    if (prm.has_sx) {
        t2i_act a = d->stat_x;
        if (a.c > y.c) a.c = y.c;      // channels beyond its extent are out of bounds for the map: zero filled, not counted
        rc = make_act_maps(a, parity_out, d->np, prm.bq, prm.bp, prm.bn, prm.sx_maps);
        if (rc != T2I_OK) return rc;
    }
    {
        const int taps = a_img ? 1 : (d->mode == T2I_CONV_S1) ? d->k * d->k : 16;
        const uint64_t e = 2;
        const uint64_t plane_bytes = (d->np == 2) ? (uint64_t)d->w_plane_stride * e : (uint64_t)taps * d->w_rows * d->w_cols * e;
        const uint64_t dims[4] = {(uint64_t)d->w_cols, (uint64_t)d->w_rows, (uint64_t)taps, (uint64_t)d->np};
        const uint64_t str[3] = {(uint64_t)d->w_cols * e, (uint64_t)d->w_cols * d->w_rows * e, plane_bytes};
        const uint32_t box[4] = {64, kn ? 64u : (uint32_t)b_rows, 1, 1};
        rc = encode_tmap_bf16(&prm.b_map, d->w, 4, dims, str, box);
        if (rc != T2I_OK) return rc;
    }
    prm.bias = d->bias;
    prm.act = d->act;
    prm.mask_kind = d->mask_kind;

    const int grid = (prm.total_tiles < workers ? prm.total_tiles : workers) * (cta2 ? 2 : 1);
    typedef void (*KernelFn)(const ConvGemmParams);
    static bool attr_done[24] = {};
    const int variant = (stats ? 12 : 0) + (cta2 ? 6 : 0) + (block_n == 256 ? 2 : block_n == 64 ? 4 : 0) + (kn ? 1 : 0);
    const KernelFn fns[24] = {
        conv_gemm_kernel<128, false, false, false>, conv_gemm_kernel<128, true, false, false>,
        conv_gemm_kernel<256, false, false, false>, conv_gemm_kernel<256, true, false, false>,
        conv_gemm_kernel<64, false, false, false>,  conv_gemm_kernel<64, true, false, false>,
        conv_gemm_kernel<128, false, true, false>,  conv_gemm_kernel<128, true, true, false>,
        conv_gemm_kernel<256, false, true, false>,  conv_gemm_kernel<256, true, true, false>,
        conv_gemm_kernel<64, false, true, false>,   nullptr,
        conv_gemm_kernel<128, false, false, true>,  conv_gemm_kernel<128, true, false, true>,
        conv_gemm_kernel<256, false, false, true>,  conv_gemm_kernel<256, true, false, true>,
        conv_gemm_kernel<64, false, false, true>,   conv_gemm_kernel<64, true, false, true>,
        conv_gemm_kernel<128, false, true, true>,   conv_gemm_kernel<128, true, true, true>,
        conv_gemm_kernel<256, false, true, true>,   conv_gemm_kernel<256, true, true, true>,
        conv_gemm_kernel<64, false, true, true>,    nullptr};
    if (a_img) {
        static bool img_attr_done[4] = {};
        const KernelFn img_fns[4] = {conv_gemm_kernel<128, false, false, false, true>, conv_gemm_kernel<128, true, false, false, true>,
                                     conv_gemm_kernel<128, false, false, true, true>, conv_gemm_kernel<128, true, false, true, true>};
        const int iv = (stats ? 2 : 0) + (kn ? 1 : 0);
        if (!img_attr_done[iv]) {
            cudaError_t e = cudaFuncSetAttribute(img_fns[iv], cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
            if (e != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            img_attr_done[iv] = true;
        }
        cudaError_t le = launch_pdl(img_fns[iv], grid, kThreads + img_producers(stats), smem_bytes, stream, prm, 1);
        if (le != cudaSuccess) return fail(T2I_ERR_CUDA, "conv_gemm_kernel (image patches) launch: %s", cudaGetErrorString(le));
        return check_launch("conv_gemm_kernel");
    }
    if (fns[variant] == nullptr) return fail(T2I_ERR_BAD_ARG, "no kernel variant %d", variant);
    if (!attr_done[variant]) {
        cudaError_t e = cudaFuncSetAttribute(fns[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_done[variant] = true;
    }
    cudaError_t le = launch_pdl(fns[variant], grid, kThreads, smem_bytes, stream, prm, cta2 ? 2 : 1);
    if (le != cudaSuccess) return fail(T2I_ERR_CUDA, "conv_gemm_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("conv_gemm_kernel");
}
