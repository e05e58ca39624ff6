// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/t2i_b200.h"

namespace t2i {

void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError after a launch -> status
int num_sms();

// Encode a bf16 tiled tensor map with 128-byte swizzle and zero OOB fill.
// dims/strides are innermost first; strides_bytes[i] is the stride of dim i+1.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

struct Tap {
    int8_t dp, dq;  // shift of the input box on the virtual pixel grid
    int8_t map;     // which input tensor map (stride-2 parity view) the tap reads
    int8_t wtap;    // tap index into the packed weights (kh*k + kw)
};
struct TapTable {
    int n_phases, taps_per_phase, n_maps;
    Tap taps[16];
    int8_t ph_op[4], ph_oq[4];  // DECONV: output sub-pixel offsets per phase
};
// Tap tables for the three conv forms (see include/t2i_b200.h); returns 0 or sets the error.
int build_taps(int mode, int k, int flip, TapTable* t);

// Launch with programmatic stream serialization (PDL): the kernel may begin while its predecessor in
// the stream drains; it must call pdl_wait() before touching global memory.
template <typename Params>
inline cudaError_t launch_pdl(void (*kernel)(const Params), int grid, int block, size_t smem, cudaStream_t stream,
                              const Params& prm, int cluster = 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    int n = 1;
    if (cluster > 1) {   // thread-block cluster (CTA pair on one TPC) for cta_group::2 kernels
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = cluster;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        n = 2;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, prm);
}

// Elementwise / reduction kernels: the same programmatic-serialization attribute, so that a kernel's launch latency
// hides under its predecessor's tail; every such kernel calls pdl_wait() before it touches global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ew(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Division by a runtime constant as multiply-high + shift (Granlund-Montgomery, round-up variant), valid for
// 0 <= n < 2^31: q = (umulhi(mul, n) + n) >> shr.  The kernels decode tile indices with it.
struct FastDiv {
    uint32_t mul, shr, d;
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (uint32_t)d;
    uint32_t l = 0;
    while ((1u << l) < (uint32_t)d) ++l;
    f.shr = l;
    f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - (uint64_t)d)) / (uint64_t)d + 1);
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
    return (int)((__umulhi(f.mul, (uint32_t)n) + (uint32_t)n) >> f.shr);
}
#endif

inline int floor_pow2(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace t2i
