#include "host_util.h"

#include <atomic>
#include <mutex>
#include <string.h>

namespace t2i {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(T2I_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    count_launch();
    return T2I_OK;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(T2I_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(T2I_ERR_BAD_ARG, "tensor base not 16B aligned");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        if (box[i] == 0 || box[i] > 256) return fail(T2I_ERR_BAD_ARG, "TMA box dim %d = %u out of range", i, box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        if (gstr[i] % 16 != 0) return fail(T2I_ERR_BAD_ARG, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                                           (unsigned long long)gstr[i]);
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                     bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(T2I_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return T2I_OK;
}

int build_taps(int mode, int k, int flip, TapTable* t) {
    memset(t, 0, sizeof(*t));
    if (mode == T2I_CONV_S1) {
        if (k < 1 || k > 4) return fail(T2I_ERR_BAD_ARG, "CONV_S1 supports k = 1 .. 4, got %d", k);
        // SAME: pad_total = k - 1, before = pad_total / 2 (k = 4: 1 before, 2 after -- TF's asymmetric case, used by
        // StackGAN stage-II's 4x4 stride-1 convs, models/stackgan/stageII/model.py:102,105,151,154)
        const int pad = (k - 1) / 2;
        t->n_phases = 1;
        t->taps_per_phase = k * k;
        t->n_maps = 1;
        for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw) {
                Tap& tp = t->taps[kh * k + kw];
                tp.dp = (int8_t)(flip ? pad - kh : kh - pad);
                tp.dq = (int8_t)(flip ? pad - kw : kw - pad);
                tp.map = 0;
                tp.wtap = (int8_t)(kh * k + kw);
            }
        return T2I_OK;
    }
    if (mode == T2I_CONV_K4S2) {
        // y[p] = sum_kh x[2p - 1 + kh] w[kh];  2p - 1 + kh = 2 (p + a) + r with kh - 1 = 2a + r
        t->n_phases = 1;
        t->taps_per_phase = 16;
        t->n_maps = 4;
        for (int kh = 0; kh < 4; ++kh)
            for (int kw = 0; kw < 4; ++kw) {
                const int rh = (kh + 1) & 1, ah = (kh - 1 - rh) / 2;
                const int rw = (kw + 1) & 1, aw = (kw - 1 - rw) / 2;
                Tap& tp = t->taps[kh * 4 + kw];
                tp.dp = (int8_t)ah;
                tp.dq = (int8_t)aw;
                tp.map = (int8_t)(rh * 2 + rw);
                tp.wtap = (int8_t)(kh * 4 + kw);
            }
        return T2I_OK;
    }
    if (mode == T2I_DECONV_K4S2) {
        // y[2p + ph] = sum over kh with (ph + 1 - kh) even of x[p + (ph + 1 - kh)/2] w[kh]
        t->n_phases = 4;
        t->taps_per_phase = 4;
        t->n_maps = 1;
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const int phase = ph * 2 + pw;
                t->ph_op[phase] = (int8_t)ph;
                t->ph_oq[phase] = (int8_t)pw;
                int idx = 0;
                for (int kh = 0; kh < 4; ++kh) {
                    if (((ph + 1 - kh) & 1) != 0) continue;
                    for (int kw = 0; kw < 4; ++kw) {
                        if (((pw + 1 - kw) & 1) != 0) continue;
                        Tap& tp = t->taps[phase * 4 + idx++];
                        tp.dp = (int8_t)((ph + 1 - kh) / 2);
                        tp.dq = (int8_t)((pw + 1 - kw) / 2);
                        tp.map = 0;
                        tp.wtap = (int8_t)(kh * 4 + kw);
                    }
                }
            }
        return T2I_OK;
    }
    return fail(T2I_ERR_BAD_ARG, "unknown conv mode %d", mode);
}

}  // namespace t2i

extern "C" const char* t2i_last_error(void) { return t2i::g_err; }
extern "C" int t2i_version(void) { return 1; }
extern "C" long long t2i_launch_count(void) { return t2i::g_launches.load(); }
