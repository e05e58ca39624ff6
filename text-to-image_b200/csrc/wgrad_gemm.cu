// Weight-gradient GEMM for sm_100a.  dw[tap][co][ci] += sum_pixels dy[pixel][co] * x[pixel+tap][ci].
// The contraction runs over pixels, which is the slow (row) index of NHWC tensors, so both
// operands are MN-major: TMA loads [64 pixels x 64 channels] boxes (128B swizzle) and the UMMA
// descriptors walk 16 pixel rows per instruction (LBO = distance between 64-channel atoms,
// SBO = 8 pixel rows).  Work items are (tap, co tile, ci tile, K split); split-K partial sums are
// combined with vectorised fp32 reductions (red.global.add.v4.f32) into dw.  The kernel is persistent:
// each CTA (pair) walks its work items with the accumulator double-buffered in TMEM, so that the reduction
// of one item and the pipeline fill of the next overlap the MMAs (one CTA per item left 7 waves of
// un-overlapped fill / drain on the split-K launches).
//
// Tile shapes (MT x 128 output channels by BN input channels): 1x128, 1x256, 2x128.  The wide
// shapes halve the operand bytes fetched per MMA cycle (48 KB per 512 tensor cycles instead of
// 32 KB per 256), which is what bounds the 128x128 shape (L2 -> SMEM feed).
//
// Replaces Conv2DBackpropFilter / MatMul-grad reached via AdamOptimizer.minimize at
// models/wgancls/model.py:94-106 of the reference; also forms the second-order term of the
// gradient penalty (model.py:62-70,88-91) when x holds the tangent activations.
#include "host_util.h"
#include "img_patch.cuh"
#include "ptx.cuh"

namespace t2i {

constexpr int kWK = 64;                       // pixels per K block
constexpr int kWAtomBytes = kWK * 128;        // 64 pixels x 64 channels bf16 = 8 KB
constexpr int kWThreads = 192;
// IMG: warps 2..7 (the four reduction warps, idle until a work item ends, plus two more) feed the patch rows of the K blocks
constexpr int kWImgProducers = 192;
constexpr int kWThreadsImg = 64 + kWImgProducers;
constexpr int kWImgRing = 8;                  // IMG: ring slots of padded bf16 image rows (bulk copies, six K blocks ahead)
constexpr int kWImgRowBytes = kWImgRing * 6400;   // a slot: (2*bp + 2) rows of one plane (the pass's), at most 6400 bytes

// CTA2: a CTA pair (cta_group::2) computes 256 output channels x BN input channels; each CTA stages its
// own 128 output channels of dy and HALF of the x tile, and reduces its own 128 accumulator rows.
template <int MT, int BN, bool CTA2 = false>
struct WgradCfg {
    static constexpr int kM = MT * 128;                       // co per CTA
    static constexpr int kBCols = CTA2 ? BN / 2 : BN;         // ci staged by this CTA
    static constexpr int kABytes = (kM / 64) * kWAtomBytes;
    static constexpr int kBBytes = (kBCols / 64) * kWAtomBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (kStageBytes <= 32768) ? 6 : 4;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kSmemBytes = kBarOffset + 256 + 1024;
    // IMG variants: four pipeline stages, then barriers, then the image-row buffers
    static constexpr int kStagesImg = 4;
    static constexpr int kBarOffsetImg = kStagesImg * kStageBytes;
    static constexpr int kImgOffset = kBarOffsetImg + 256;
    static constexpr int kSmemBytesImg = kImgOffset + kWImgRowBytes + 1024;
    static constexpr int kAccCols = MT * BN;                  // 128 or 256 columns per accumulator
    static constexpr int kTmemCols = 2 * MT * BN;             // double buffered
};

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct alignas(64) WgradParams {
    CUtensorMap x_maps[4];
    CUtensorMap dy_maps[4];
    TapTable tt;
    int N, P, Q;
    int bn, bp, bq;        // K box, bn*bp*bq == 64
    int tiles_p, tiles_q, k_blocks;  // k_blocks = pixel boxes covering the virtual grid
    int tiles_co, tiles_ci, jobs;    // jobs = n_phases * taps_per_phase
    int splits, kb_per_split;
    int total_work;        // jobs * tiles_co * tiles_ci * splits
    int tap_fast;          // work order: taps fastest (the taps of one pixel range run together and share x / dy in L2)
    int n_pass;
    int cout, cin;
    float* dw;
    // IMG: one operand is the 4x4 / stride-2 patch matrix of this fp32 NHWC 3-channel image (64 columns, 48 used),
    // assembled in shared memory (see conv_gemm.cu, A_IMG)
    const uint32_t* img;       // padded bf16 rows [np][N][img_h][pitch] as words (t2i_img_to_rows)
    long long img_plane_words;
    int img_h, img_w;
    unsigned long long* timeline;      // debug build only (T2I_TIMELINE_BUILD), see conv_gemm.cu
};
#ifdef T2I_TIMELINE_BUILD
#define T2I_WMARK(it, ev) \
    do { if (prm.timeline != nullptr && blockIdx.x == 0 && (it) < 64) prm.timeline[(it) * 8 + (ev)] = global_timer_ns(); } while (0)
#else
#define T2I_WMARK(it, ev) do {} while (0)
#endif

// IMG = 1: the x operand (B) is the image patch matrix -- d_net's first conv: dw[co][(kh,kw,c)] += dy[pixel][co] * patch;
// IMG = 2: the dy operand (A) is -- g_net's last transposed conv: dw[(kh,kw,c)][ci] += patch(d image)[pixel] * x[pixel][ci].
template <int MT, int BN, bool CTA2, int IMG = 0>
__global__ void __launch_bounds__(IMG ? kWThreadsImg : kWThreads, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams prm) {
    using Cfg = WgradCfg<MT, BN, CTA2>;
    static_assert(!CTA2 || MT == 1, "the CTA-pair variant uses one accumulator per CTA");
    static_assert(IMG == 0 || (!CTA2 && MT == 1), "image-patch operands: single CTA, one accumulator");
    constexpr int kPair = CTA2 ? 2 : 1;
    const int rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;
    constexpr int kStages = IMG ? Cfg::kStagesImg : Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (IMG ? Cfg::kBarOffsetImg : Cfg::kBarOffset));
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* rows_full = tmem_empty + 2;       // IMG: one per ring slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rows_full + kWImgRing);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < prm.tt.n_maps; ++i) tma_prefetch_desc(&prm.x_maps[i]);
        if (IMG != 2)      // IMG = 2: dy is the image patch matrix, there is no dy tensor map
            for (int i = 0; i < (prm.tt.n_phases > 1 ? prm.tt.n_phases : 1); ++i) tma_prefetch_desc(&prm.dy_maps[i]);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], IMG ? 1 + kWK : kPair);      // IMG: the TMA thread + the 64 threads that write patch rows
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kPair * 4);      // the four reduction warps of both CTAs (leader's copy)
        }
        if (IMG != 0)
            for (int i = 0; i < kWImgRing; ++i) mbar_init(&rows_full[i], 1);
        fence_barrier_init();
    }
    if (CTA2) cluster_sync_all();
    if (warp == 1) {
        if (CTA2) {
            tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, Cfg::kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // work decode: split fastest so that the CTAs of one output tile run together
    struct Work {
        int split, cit, cot, job, kb_begin, kb_end, n_kb;
    };
    auto decode = [&](int w) {
        Work k;
        if (prm.tap_fast) {   // operands larger than L2: every (tap, channel tile) item of a pixel range before the next range
            k.job = w % prm.jobs; w /= prm.jobs;
            k.cit = w % prm.tiles_ci; w /= prm.tiles_ci;
            k.cot = w % prm.tiles_co; w /= prm.tiles_co;
            k.split = w;
        } else {
            k.split = w % prm.splits; w /= prm.splits;
            k.cit = w % prm.tiles_ci; w /= prm.tiles_ci;
            k.cot = w % prm.tiles_co; w /= prm.tiles_co;
            k.job = w;   // phase * taps_per_phase + t
        }
        k.kb_begin = k.split * prm.kb_per_split;
        k.kb_end = k.kb_begin + prm.kb_per_split;
        if (k.kb_end > prm.k_blocks) k.kb_end = prm.k_blocks;
        k.n_kb = (k.kb_end > k.kb_begin) ? (k.kb_end - k.kb_begin) * prm.n_pass : 0;
        return k;
    };
    const int w0 = blockIdx.x / kPair, w_stride = gridDim.x / kPair;

    // accumulator of work item wk (the it-th of this CTA) -> red.global.add into dw; executed by warps 2..5
    auto reduce_item = [&](const Work& wk, int it) {
        const int quarter = warp & 3;
        const int acc = it & 1;
        const Tap tap = prm.tt.taps[wk.job];
        const int co_cta = (wk.cot * kPair + rank) * Cfg::kM;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1, 400 + acc);
        tc_fence_after();
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const int co = co_cta + mt * 128 + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + acc * Cfg::kAccCols + mt * BN + (static_cast<uint32_t>(quarter * 32) << 16);
            float* row = prm.dw + (static_cast<long long>(tap.wtap) * prm.cout + co) * prm.cin;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int ci0 = wk.cit * BN + c0;
                if (ci0 >= prm.cin) break;
                __syncwarp();
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                if (co < prm.cout) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int ci = ci0 + g * 4;
                        if (ci < prm.cin)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + ci),
                                         "f"(__uint_as_float(r[g * 4 + 0])), "f"(__uint_as_float(r[g * 4 + 1])),
                                         "f"(__uint_as_float(r[g * 4 + 2])), "f"(__uint_as_float(r[g * 4 + 3]))
                                         : "memory");
                    }
                }
            }
        }
        // accumulator read: hand the TMEM buffer back to the MMA warp (of the leader CTA)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (CTA2 && rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
            else mbar_arrive(&tmem_empty[acc]);
        }
    };

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = w0; w < prm.total_work; w += w_stride) {
                const Work wk = decode(w);
                const Tap tap = prm.tt.taps[wk.job];
                const CUtensorMap* dy_map = &prm.dy_maps[prm.tt.n_phases > 1 ? wk.job / prm.tt.taps_per_phase : 0];
                const CUtensorMap* x_map = &prm.x_maps[tap.map];
                const int co_cta = (wk.cot * kPair + rank) * Cfg::kM;        // first output channel of this CTA
                const int ci_cta = wk.cit * BN + rank * Cfg::kBCols;         // first input channel staged by this CTA
                int a_atoms = (prm.cout - co_cta + 63) / 64;                 // atoms of dy with at least one real channel
                if (a_atoms > Cfg::kM / 64) a_atoms = Cfg::kM / 64;
                if (a_atoms < 1) a_atoms = 1;
                for (int pass = 0; pass < prm.n_pass; ++pass) {
                    const int pa = (pass == 1) ? 1 : 0;  // dy plane: hi, lo, hi
                    const int pb = (pass == 2) ? 1 : 0;  // x  plane: hi, hi, lo
                    for (int kb = wk.kb_begin; kb < wk.kb_end; ++kb) {
                        const int tq = kb % prm.tiles_q;
                        const int tp = (kb / prm.tiles_q) % prm.tiles_p;
                        const int tn = kb / (prm.tiles_q * prm.tiles_p);
                        const int q0 = tq * prm.bq, p0 = tp * prm.bp, n0 = tn * prm.bn;
                        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                        uint8_t* sa = smem + stage * Cfg::kStageBytes;
                        if (IMG == 1) {            // dy by TMA; the patch atom (B) comes from the producer warps
                            mbar_arrive_expect_tx(&full_bar[stage], a_atoms * kWAtomBytes);
#pragma unroll
                            for (int a = 0; a < Cfg::kM / 64; ++a)
                                if (a < a_atoms)
                                    tma_load_5d(dy_map, &full_bar[stage], sa + a * kWAtomBytes, co_cta + a * 64, q0, p0, n0, pa);
                        } else if (IMG == 2) {     // x by TMA; the patch atom (A, 64 of the 128 rows) from the producer warps
                            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
#pragma unroll
                            for (int b = 0; b < Cfg::kBCols / 64; ++b)
                                tma_load_5d(x_map, &full_bar[stage], sa + Cfg::kABytes + b * kWAtomBytes, ci_cta + b * 64,
                                            q0 + tap.dq, p0 + tap.dp, n0, pb);
                        } else if (!CTA2) {
                            // 64-channel atoms of dy that lie entirely beyond cout are not fetched: their accumulator
                            // rows are never written out, so whatever the shared memory holds there is harmless
                            mbar_arrive_expect_tx(&full_bar[stage], a_atoms * kWAtomBytes + Cfg::kBBytes);
#pragma unroll
                            for (int a = 0; a < Cfg::kM / 64; ++a)
                                if (a < a_atoms)
                                    tma_load_5d(dy_map, &full_bar[stage], sa + a * kWAtomBytes, co_cta + a * 64, q0, p0, n0, pa);
#pragma unroll
                            for (int b = 0; b < Cfg::kBCols / 64; ++b)
                                tma_load_5d(x_map, &full_bar[stage], sa + Cfg::kABytes + b * kWAtomBytes, ci_cta + b * 64,
                                            q0 + tap.dq, p0 + tap.dp, n0, pb);
                        } else {
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                            else mbar_arrive_cluster(&full_bar[stage], 0);
#pragma unroll
                            for (int a = 0; a < Cfg::kM / 64; ++a)
                                tma_load_5d_2sm(dy_map, &full_bar[stage], sa + a * kWAtomBytes, co_cta + a * 64, q0, p0, n0, pa);
#pragma unroll
                            for (int b = 0; b < Cfg::kBCols / 64; ++b)
                                tma_load_5d_2sm(x_map, &full_bar[stage], sa + Cfg::kABytes + b * kWAtomBytes, ci_cta + b * 64,
                                                q0 + tap.dq, p0 + tap.dp, n0, pb);
                        }
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kPair * 128, BN, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int w = w0; w < prm.total_work; w += w_stride) {
                const Work wk = decode(w);
                if (wk.n_kb == 0) continue;
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, 300 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
                for (int kb = 0; kb < wk.n_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 200 + stage);
                    tc_fence_after();
                    if (IMG != 0 && (kb & 1) == 0) T2I_WMARK(kb >> 1, 5);
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
#pragma unroll
                    for (int k = 0; k < kWK / 16; ++k) {
                        // 16 pixel rows of 128 B per instruction
                        const uint64_t db = make_sw128_desc(sa + Cfg::kABytes + k * 2048, kWAtomBytes, 1024);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint64_t da = make_sw128_desc(sa + mt * 2 * kWAtomBytes + k * 2048, kWAtomBytes, 1024);
                            if (CTA2) umma_bf16_2sm(d_tmem + mt * BN, da, db, idesc, (kb | k) != 0);
                            else umma_bf16(d_tmem + mt * BN, da, db, idesc, (kb | k) != 0);
                        }
                    }
                    if (CTA2) {
                        umma_commit_2sm(&empty_bar[stage]);
                        if (kb == wk.n_kb - 1) umma_commit_2sm(&tmem_full[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == wk.n_kb - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ++it;
            }
        }
    } else if (IMG != 0) {
        // warps 2..7: image-patch producers (one pixel row of the K block per thread for the first 64 of them); warps
        // 2..5 reduce a work item when its last block has been produced.  The K blocks this CTA will visit form one
        // sequence (work item, pass, block); the padded bf16 image rows of the next kWImgRing - 1 blocks are in flight.
        const int pt = threadIdx.x - 64;
        uint32_t* s_img = reinterpret_cast<uint32_t*>(smem + Cfg::kImgOffset);
        const int n_rows = 2 * prm.bp + 2;
        const int pitchw = img_pitch_words(prm.img_w);
        const int slot_words = n_rows * pitchw;
        const long long sample_words = static_cast<long long>(prm.img_h) * pitchw;
        // (the block coordinates travel with the cursor: the producers are few and latency bound, integer divisions in
        // this loop cost more than the patch assembly itself -- measured with the debug timeline)
        struct Cursor {
            int w, pass, kb, kb_begin, kb_end;
            int tq, tp, tn, tq0, tp0, tn0;        // block coordinates of kb and of kb_begin (bn == 1: tn = the sample)
            bool valid;
        };
        auto settle = [&](Cursor& c) {      // move to the first existing block at or after (w, pass = 0)
            while (c.w < prm.total_work) {
                const Work wk = decode(c.w);
                if (wk.n_kb > 0) {
                    c.kb_begin = wk.kb_begin; c.kb_end = wk.kb_end; c.kb = wk.kb_begin; c.pass = 0;
                    c.tq0 = c.kb % prm.tiles_q;
                    c.tp0 = (c.kb / prm.tiles_q) % prm.tiles_p;
                    c.tn0 = c.kb / (prm.tiles_q * prm.tiles_p);
                    c.tq = c.tq0; c.tp = c.tp0; c.tn = c.tn0;
                    c.valid = true;
                    return;
                }
                c.w += w_stride;
            }
            c.valid = false;
        };
        auto advance = [&](Cursor& c) {
            if (++c.kb < c.kb_end) {
                if (++c.tq == prm.tiles_q) {
                    c.tq = 0;
                    if (++c.tp == prm.tiles_p) {
                        c.tp = 0;
                        ++c.tn;
                    }
                }
                return;
            }
            if (++c.pass < prm.n_pass) {
                c.kb = c.kb_begin;
                c.tq = c.tq0; c.tp = c.tp0; c.tn = c.tn0;
                return;
            }
            c.w += w_stride;
            settle(c);
        };
        const int lg_bq = __ffs(prm.bq) - 1;
        // `who`: the 64-thread group that issues this fetch (two fetches of a trip run side by side in two warps)
        auto fetch = [&](const Cursor& c, int slot, int who) {
            if (!c.valid) return;
            const int fpt = pt - 64 * who;
            if (fpt < 0 || fpt >= 64) return;
            const int tp = c.tp, tn = c.tn;
            const bool lo_plane = (IMG == 1) ? (c.pass == 2) : (c.pass == 1);     // x: hi, hi, lo; dy: hi, lo, hi
            img_rows_fetch(prm.img + (lo_plane ? prm.img_plane_words : 0), 0, sample_words, 1, tn, 2 * tp * prm.bp - 1,
                           prm.img_h, n_rows, pitchw, s_img + slot * slot_words, &rows_full[slot], fpt, 64);
        };
        // Two K blocks per trip: threads 0..63 assemble block j, threads 64..127 block j + 1 (one named barrier and one
        // round of bulk-copy issue per 128 pixels); the ring holds kWImgRing slots, kWImgRing - 2 blocks are in flight.
        Cursor cur;
        cur.w = w0;
        settle(cur);
        Cursor pf = cur;                         // runs kWImgRing - 2 blocks ahead of cur
        for (int a = 0; a < kWImgRing - 2; ++a) {
            fetch(pf, a, a & 1);
            if (pf.valid) advance(pf);
        }
        int jb0 = 0, it = 0;                     // jb0: index of `cur` in this CTA's block sequence
        const int gi = pt >> 6, lt = pt & 63;
        while (cur.valid) {
            Cursor c1 = cur;
            advance(c1);                         // second block of the trip (may not exist)
            Cursor c2 = c1;
            if (c1.valid) advance(c2);           // first block of the next trip
            named_bar(3, kWImgProducers);        // everybody is done with the previous trip: its two ring slots are free
            if (pt == 0) T2I_WMARK(jb0 >> 1, 0);
            fetch(pf, (jb0 + kWImgRing - 2) & (kWImgRing - 1), 0);
            if (pf.valid) advance(pf);
            fetch(pf, (jb0 + kWImgRing - 1) & (kWImgRing - 1), 1);
            if (pf.valid) advance(pf);
            if (gi < 2) {
                const Cursor& cb = (gi == 0) ? cur : c1;
                if (cb.valid) {
                    const int jb = jb0 + gi;
                    const int slot = jb & (kWImgRing - 1), stage = jb % kStages;
                    if (pt == 0) T2I_WMARK(jb0 >> 1, 1);
                    mbar_wait(&rows_full[slot], (jb / kWImgRing) & 1, 800 + slot);
                    if (pt == 0) T2I_WMARK(jb0 >> 1, 2);
                    const int q0 = cb.tq * prm.bq;
                    mbar_wait(&empty_bar[stage], ((jb / kStages) & 1) ^ 1, 700 + stage);
                    if (pt == 0) T2I_WMARK(jb0 >> 1, 3);
                    uint8_t* atom = smem + stage * Cfg::kStageBytes + (IMG == 1 ? Cfg::kABytes : 0);
                    int q = q0 + (lt & (prm.bq - 1));
                    if (q >= prm.Q) q = prm.Q - 1;
                    img_patch_row(s_img + slot * slot_words, pitchw, lt >> lg_bq, q, atom + (lt >> 3) * 1024 + (lt & 7) * 128, lt,
                                  true);   // columns 48..63: zeros
                    fence_proxy_async();
                    mbar_arrive(&full_bar[stage]);
                    if (pt == 0) T2I_WMARK(jb0 >> 1, 4);
                }
            }
            // work items completed by this trip, in order
            if (!c1.valid || c1.w != cur.w) {
                if (warp < 6) reduce_item(decode(cur.w), it);
                ++it;
            }
            if (c1.valid && (!c2.valid || c2.w != c1.w)) {
                if (warp < 6) reduce_item(decode(c1.w), it);
                ++it;
            }
            jb0 += 2;
            cur = c1.valid ? c2 : c1;
        }
    } else {
        // reduction warps: accumulator -> red.global.add into dw, while the MMAs of the next item run
        int it = 0;
        for (int w = w0; w < prm.total_work; w += w_stride) {
            const Work wk = decode(w);
            if (wk.n_kb == 0) continue;
            reduce_item(wk, it);
            ++it;
        }
    }

    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

static int make_maps(const t2i_act& t, bool parity, int np, int bq, int bp, int bn, CUtensorMap* maps) {
    if (t.pitch % 8 != 0 || t.coff % 8 != 0 || t.c % 8 != 0)
        return fail(T2I_ERR_BAD_ARG, "activation channels must be multiples of 8 (c=%d pitch=%d coff=%d)", t.c, t.pitch,
                    t.coff);
    const uint64_t e = 2;
    const uint64_t plane_bytes = (np == 2) ? (uint64_t)t.plane_stride * e : (uint64_t)t.n * t.h * t.w * t.pitch * e;
    const uint32_t box[5] = {64, (uint32_t)bq, (uint32_t)bp, (uint32_t)bn, 1};
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(t.ptr) + t.coff;
    if (!parity) {
        const uint64_t dims[5] = {(uint64_t)t.c, (uint64_t)t.w, (uint64_t)t.h, (uint64_t)t.n, (uint64_t)np};
        const uint64_t str[4] = {(uint64_t)t.pitch * e, (uint64_t)t.w * t.pitch * e, (uint64_t)t.h * t.w * t.pitch * e,
                                 plane_bytes};
        return encode_tmap_bf16(&maps[0], base, 5, dims, str, box);
    }
    if ((t.h & 1) || (t.w & 1)) return fail(T2I_ERR_BAD_ARG, "stride-2 views need even h, w (got %d x %d)", t.h, t.w);
    for (int rh = 0; rh < 2; ++rh)
        for (int rw = 0; rw < 2; ++rw) {
            const uint64_t dims[5] = {(uint64_t)t.c, (uint64_t)t.w / 2, (uint64_t)t.h / 2, (uint64_t)t.n, (uint64_t)np};
            const uint64_t str[4] = {2 * (uint64_t)t.pitch * e, 2 * (uint64_t)t.w * t.pitch * e,
                                     (uint64_t)t.h * t.w * t.pitch * e, plane_bytes};
            int rc = encode_tmap_bf16(&maps[rh * 2 + rw], base + ((long long)rh * t.w + rw) * t.pitch, 5, dims, str, box);
            if (rc != T2I_OK) return rc;
        }
    return T2I_OK;
}

template <int MT, int BN, bool CTA2>
static int launch_wgrad(const WgradParams& prm, int grid, cudaStream_t stream) {
    using Cfg = WgradCfg<MT, BN, CTA2>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel<MT, BN, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    cudaError_t le = launch_pdl(wgrad_gemm_kernel<MT, BN, CTA2>, grid * (CTA2 ? 2 : 1), kWThreads, Cfg::kSmemBytes, stream,
                                prm, CTA2 ? 2 : 1);
    if (le != cudaSuccess) return fail(T2I_ERR_CUDA, "wgrad_gemm_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("wgrad_gemm_kernel");
}

template <int BN, int IMG>
static int launch_wgrad_img(const WgradParams& prm, int grid, cudaStream_t stream) {
    using Cfg = WgradCfg<1, BN, false>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel<1, BN, false, IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytesImg);
        if (e != cudaSuccess) return fail(T2I_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    cudaError_t le = launch_pdl(wgrad_gemm_kernel<1, BN, false, IMG>, grid, kWThreadsImg, Cfg::kSmemBytesImg, stream, prm, 1);
    if (le != cudaSuccess) return fail(T2I_ERR_CUDA, "wgrad_gemm_kernel (image patches) launch: %s", cudaGetErrorString(le));
    return check_launch("wgrad_gemm_kernel");
}

}  // namespace t2i

using namespace t2i;

#ifdef T2I_TIMELINE_BUILD
unsigned long long* g_wgrad_timeline = nullptr;
extern "C" int t2i_debug_wgrad_timeline(unsigned long long* host_dst, int count) {
    if (g_wgrad_timeline == nullptr) return -1;
    return cudaMemcpy(host_dst, g_wgrad_timeline, 8 * count, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
#endif
extern "C" int t2i_wgrad_img(const void* img, long long plane_stride, int n, int h, int w, const t2i_act* other, int img_side,
                             int np, float* dw, int cout, int cin, void* stream_) {
    if ((reinterpret_cast<uintptr_t>(img) & 15) != 0) return fail(T2I_ERR_BAD_ARG, "wgrad_img: rows must be 16-byte aligned");
    if (img == nullptr || other == nullptr || other->ptr == nullptr || dw == nullptr) return fail(T2I_ERR_BAD_ARG, "null tensor");
    if (np != 1 && np != 2) return fail(T2I_ERR_BAD_ARG, "np must be 1 or 2");
    if (img_side != 1 && img_side != 2) return fail(T2I_ERR_BAD_ARG, "img_side must be 1 (x) or 2 (dy)");
    if ((h & 1) || (w & 3)) return fail(T2I_ERR_BAD_ARG, "wgrad_img: image %d x %d (need even height, width multiple of 4)", h, w);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WgradParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.tt.n_phases = 1; prm.tt.taps_per_phase = 1; prm.tt.n_maps = 0;
    prm.N = n; prm.P = h / 2; prm.Q = w / 2;
    if (other->n != n || other->h != prm.P || other->w != prm.Q)
        return fail(T2I_ERR_BAD_ARG, "wgrad_img: tensor [%d,%d,%d] does not sit on the %d x %d x %d patch grid", other->n, other->h,
                    other->w, n, prm.P, prm.Q);
    {
        int q = floor_pow2(prm.Q); if (q > kWK) q = kWK;
        int p = floor_pow2(prm.P); if (p > kWK / q) p = kWK / q;
        prm.bq = q; prm.bp = p; prm.bn = kWK / (q * p);
    }
    if (prm.bn != 1 || (2 * prm.bp + 2) * img_pitch_words(w) * 4 > kWImgRowBytes / kWImgRing)
        return fail(T2I_ERR_BAD_ARG, "wgrad_img: unsupported image extent %d x %d", h, w);
    prm.tiles_q = ceil_div(prm.Q, prm.bq);
    prm.tiles_p = ceil_div(prm.P, prm.bp);
    prm.k_blocks = n * prm.tiles_p * prm.tiles_q;
    prm.cout = cout;
    prm.cin = cin;
    int bnn = 64;
    if (img_side == 1) {
        if (cin != 64 || other->c > cout) return fail(T2I_ERR_BAD_ARG, "wgrad_img: dw must be [cout >= %d][64] (got [%d][%d])", other->c, cout, cin);
        prm.tiles_co = ceil_div(other->c, 128);
        prm.tiles_ci = 1;
    } else {
        if (cout != 64 || other->c > cin || cin % 4 != 0)
            return fail(T2I_ERR_BAD_ARG, "wgrad_img: dw must be [64][cin >= %d] (got [%d][%d])", other->c, cout, cin);
        bnn = other->c <= 64 ? 64 : 128;
        prm.tiles_co = 1;
        prm.tiles_ci = ceil_div(other->c, bnn);
    }
    prm.jobs = 1;
    prm.n_pass = (np == 2) ? 3 : 1;
    const int tiles = prm.tiles_co * prm.tiles_ci;
    const int workers = num_sms();
    int splits = workers / tiles;
    if (splits > prm.k_blocks / 8) splits = prm.k_blocks / 8;
    if (splits < 1) splits = 1;
    prm.kb_per_split = ceil_div(prm.k_blocks, splits);
    prm.splits = ceil_div(prm.k_blocks, prm.kb_per_split);
    prm.dw = dw;
#ifdef T2I_TIMELINE_BUILD
    {
        static unsigned long long* tl = nullptr;
        if (tl == nullptr) { cudaMalloc(&tl, 64 * 8 * 8); cudaMemset(tl, 0, 64 * 8 * 8); }
        prm.timeline = tl;
        extern unsigned long long* g_wgrad_timeline;
        g_wgrad_timeline = tl;
    }
#endif
    prm.img = static_cast<const uint32_t*>(img); prm.img_plane_words = plane_stride / 2; prm.img_h = h; prm.img_w = w;
    int rc = make_maps(*other, false, np, prm.bq, prm.bp, prm.bn, img_side == 1 ? prm.dy_maps : prm.x_maps);
    if (rc != T2I_OK) return rc;
    prm.total_work = tiles * prm.splits;
    const int grid = prm.total_work < workers ? prm.total_work : workers;
    if (img_side == 1) return launch_wgrad_img<64, 1>(prm, grid, stream);
    if (bnn == 64) return launch_wgrad_img<64, 2>(prm, grid, stream);
    return launch_wgrad_img<128, 2>(prm, grid, stream);
}

extern "C" int t2i_wgrad_gemm(const t2i_wgrad_desc* d, void* stream_) {
    if (d == nullptr) return fail(T2I_ERR_BAD_ARG, "null descriptor");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WgradParams prm;
    memset(&prm, 0, sizeof(prm));
    int rc = build_taps(d->mode, d->k, 0, &prm.tt);
    if (rc != T2I_OK) return rc;
    if (d->np != 1 && d->np != 2) return fail(T2I_ERR_BAD_ARG, "np must be 1 or 2");
    const t2i_act& x = d->x;
    const t2i_act& dy = d->dy;
    if (x.ptr == nullptr || dy.ptr == nullptr || d->dw == nullptr) return fail(T2I_ERR_BAD_ARG, "null tensor");
    if (d->cin % 4 != 0 || x.c > d->cin || dy.c > d->cout)
        return fail(T2I_ERR_BAD_ARG, "bad channels x.c=%d cin=%d dy.c=%d cout=%d", x.c, d->cin, dy.c, d->cout);
    int eh, ew;  // expected dy extent
    if (d->mode == T2I_CONV_S1) {
        prm.P = x.h; prm.Q = x.w; eh = x.h; ew = x.w;
    } else if (d->mode == T2I_CONV_K4S2) {
        prm.P = x.h / 2; prm.Q = x.w / 2; eh = x.h / 2; ew = x.w / 2;
    } else {
        prm.P = x.h; prm.Q = x.w; eh = 2 * x.h; ew = 2 * x.w;
    }
    prm.N = x.n;
    if (dy.n != x.n || dy.h != eh || dy.w != ew)
        return fail(T2I_ERR_BAD_ARG, "dy shape [%d,%d,%d] does not match expected [%d,%d,%d]", dy.n, dy.h, dy.w, x.n, eh, ew);
    {
        int q = floor_pow2(prm.Q); if (q > kWK) q = kWK;
        int p = floor_pow2(prm.P); if (p > kWK / q) p = kWK / q;
        prm.bq = q; prm.bp = p; prm.bn = kWK / (q * p);
    }
    prm.tiles_q = ceil_div(prm.Q, prm.bq);
    prm.tiles_p = ceil_div(prm.P, prm.bp);
    prm.k_blocks = ceil_div(prm.N, prm.bn) * prm.tiles_p * prm.tiles_q;
    prm.cout = d->cout;
    prm.cin = d->cin;
    // tile shape: prefer 256 input channels per tile, else 256 output channels (two accumulators); layers with at
    // most 64 input channels (the 128x128 / 256x256 maps of StackGAN stage-II and PGGAN) take a 64-wide tile
    int mt = 1, bnn = 128;
    if (x.c >= 256) bnn = 256;
    else if (dy.c >= 256) mt = 2;
    else if (x.c <= 64) bnn = 64;
    static const bool allow_cta2 = [] { const char* e = getenv("T2I_WGRAD_CTA2"); return !(e && e[0] == '0'); }();
    const bool cta2 = allow_cta2 && bnn == 256 && dy.c >= 256;      // CTA pair: 256 co x 256 ci
    if (cta2) mt = 2;                                               // tile covers 256 output channels (128 per CTA)
    prm.tiles_co = ceil_div(dy.c, mt * 128);
    prm.tiles_ci = ceil_div(x.c, bnn);
    prm.jobs = prm.tt.n_phases * prm.tt.taps_per_phase;
    prm.n_pass = (d->np == 2) ? 3 : 1;
    const int tiles = prm.jobs * prm.tiles_co * prm.tiles_ci;
    // Several taps and operands that cannot stay in L2 between them (each (tap, channel tile) item is a pass over its x /
    // dy slices): order the work PIXEL-RANGE-MAJOR -- all items of a range before the next range -- so that a range is
    // fetched from HBM once and shared through L2.  (Split-fastest order re-read the 4B-batch operands of d_net's 4x4 /
    // stride-2 layers 2.5-2.7 times: 570 MB of DRAM traffic for 210 MB of tensors, profiles/r02_traffic.json.)
    const long long operand_bytes = (long long)prm.k_blocks * kWK * (x.c + dy.c) * 2 * d->np;
    const int chan_tiles = prm.tiles_co * prm.tiles_ci;
    static const bool allow_range_major = [] { const char* e = getenv("T2I_WGRAD_RANGE_MAJOR"); return !(e && e[0] == '0'); }();
    prm.tap_fast = (prm.jobs > 1 && operand_bytes > (64ll << 20) && (chan_tiles == 1 || allow_range_major)) ? 1 : 0;
    int splits = d->split_k;
    if (splits <= 0) {
        // Cost model: the launch runs in ceil(tiles * s / workers) waves; one wave costs the K blocks of a split
        // plus a fixed prologue/epilogue (about 10 K-block times).  Pick the split count with the least total.
        const int workers = cta2 ? num_sms() / 2 : num_sms();
        const int max_splits = prm.k_blocks / 8 > 0 ? prm.k_blocks / 8 : 1;  // >= 8 K blocks per CTA
        long long best = -1;
        splits = 1;
        for (int sct = 1; sct <= max_splits && sct <= 64; ++sct) {
            const long long waves = ceil_div(tiles * sct, workers);
            const long long cost = waves * (ceil_div(prm.k_blocks, sct) * prm.n_pass + 10);
            if (best < 0 || cost < best) {
                best = cost;
                splits = sct;
            }
        }
    }
    if (prm.tap_fast && d->split_k <= 0 && chan_tiles == 1) {
        // short pixel ranges (~96 K blocks = 6144 pixels) so that the taps of a range, launched side by side, stay
        // within L2 reach of each other; round the item count up to whole waves
        const int workers = num_sms();
        long long sct = prm.k_blocks / 96;
        if (sct < splits) sct = splits;
        if (sct > 4096) sct = 4096;
        const long long waves = ceil_div((int)(sct * prm.jobs), workers);
        sct = waves * workers / prm.jobs;
        if (sct >= 1) splits = (int)sct;
    } else if (prm.tap_fast && d->split_k <= 0) {
        // several channel tiles: the cost model's split count stands (it has counted the waves); only if a range's
        // operands would not stay in L2 while its items run (a few waves) are the ranges made shorter (<= ~48 MB each)
        long long sct = ceil_div((int)(operand_bytes >> 20), 48);
        const long long cap = prm.k_blocks / 16 > 0 ? prm.k_blocks / 16 : 1;
        if (sct > cap) sct = cap;
        if (sct > splits) splits = (int)sct;
    }
    if (splits > prm.k_blocks) splits = prm.k_blocks;
    prm.kb_per_split = ceil_div(prm.k_blocks, splits);
    prm.splits = ceil_div(prm.k_blocks, prm.kb_per_split);
    prm.dw = d->dw;

    rc = make_maps(x, d->mode == T2I_CONV_K4S2, d->np, prm.bq, prm.bp, prm.bn, prm.x_maps);
    if (rc != T2I_OK) return rc;
    rc = make_maps(dy, d->mode == T2I_DECONV_K4S2, d->np, prm.bq, prm.bp, prm.bn, prm.dy_maps);
    if (rc != T2I_OK) return rc;

    prm.total_work = tiles * prm.splits;
    const int workers_all = cta2 ? num_sms() / 2 : num_sms();
    if (prm.total_work <= workers_all) prm.tap_fast = 0;       // a single wave: the order does not matter
    const int grid = prm.total_work < workers_all ? prm.total_work : workers_all;     // persistent CTAs (pairs)
    if (cta2) return launch_wgrad<1, 256, true>(prm, grid, stream);
    if (bnn == 256) return launch_wgrad<1, 256, false>(prm, grid, stream);
    if (mt == 2) return launch_wgrad<2, 128, false>(prm, grid, stream);
    if (bnn == 64) return launch_wgrad<1, 64, false>(prm, grid, stream);
    return launch_wgrad<1, 128, false>(prm, grid, stream);
}
