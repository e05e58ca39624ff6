// bf16-plane storage helpers shared by the HBM-bound kernels (elementwise.cu, pggan_ops.cu): 16-byte vector access to
// [np][...] planes (np = 1: bf16; np = 2: hi + lo split), warp / block reductions, grid sizing.
#pragma once
#include "host_util.h"
#include "ptx.cuh"

namespace t2i {

typedef __nv_bfloat16 bf16;

// ---- bf16 planes: 8 consecutive values (16 bytes per plane) --------------------------------
__device__ __forceinline__ void unpack8(const uint4 u, float* v) {
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void load8(const bf16* p, long long ps, int np, float* v) {
    unpack8(*reinterpret_cast<const uint4*>(p), v);
    if (np == 2) {
        float l[8];
        unpack8(*reinterpret_cast<const uint4*>(p + ps), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += l[j];
    }
}
__device__ __forceinline__ void store8(bf16* p, long long ps, int np, const float* v) {
    uint4 hi;
    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
    hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = hi;
    if (np == 2) {
        float h[8];
        unpack8(hi, h);
        uint4 lo;
        lo.x = pack_bf16x2(v[0] - h[0], v[1] - h[1]); lo.y = pack_bf16x2(v[2] - h[2], v[3] - h[3]);
        lo.z = pack_bf16x2(v[4] - h[4], v[5] - h[5]); lo.w = pack_bf16x2(v[6] - h[6], v[7] - h[7]);
        *reinterpret_cast<uint4*>(p + ps) = lo;
    }
}
// 8 consecutive fp32 values (32-byte aligned)
__device__ __forceinline__ void load_f8(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ float load1(const bf16* p, long long ps, int np) {
    float v = __bfloat162float(p[0]);
    if (np == 2) v += __bfloat162float(p[ps]);
    return v;
}
__device__ __forceinline__ void store1(bf16* p, long long ps, int np, float v) {
    const bf16 h = __float2bfloat16_rn(v);
    p[0] = h;
    if (np == 2) p[ps] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of one float per thread; result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;
}

static inline int grid_for(long long work_items, int threads, int max_blocks_per_sm = 8) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * max_blocks_per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace t2i
