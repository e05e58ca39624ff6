// HBM-bound kernels that the conditional progressive-growing WGAN (models/pggan/pggan.py of the reference) adds to
// the shared set: per-sample layer normalisation (utils/ops.py:74-81) forward / backward, 2x nearest-neighbour
// upscale and 2x2 average pool (utils/ops.py:100-111) and their transposes, the fade-in blend (pggan.py:268,313) and the
// 3 <-> 8 channel padding of the image ends.  Same conventions as elementwise.cu: bf16 planes, 16-byte vector access,
// fp32 arithmetic, programmatic stream serialization.
#include "planes.cuh"

namespace t2i {

// per-sample mean / rstd from [sum x | sum x^2] over m values (biased variance, as tf.nn.moments)
__device__ __forceinline__ void ln_mean_rstd(const float* __restrict__ sums, int smp, float inv_m, float eps, float& mean,
                                             float& rstd) {
    const float s = sums[2 * smp], q = sums[2 * smp + 1];
    mean = s * inv_m;
    const float var = fmaxf(q * inv_m - mean * mean, 0.f);
    rstd = rsqrtf(var + eps);
}

// ---- layer_norm forward ---------------------------------------------------------------------------------------
// sums[n][2] += [sum x | sum x^2] of sample n (grid.y = sample, grid.x = chunk of the sample's m / 8 vectors)
__global__ void ln_stats_kernel(const bf16* __restrict__ x, long long ps, int np, long long m8, float* sums) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    const int smp = blockIdx.y;
    const bf16* base = x + (long long)smp * m8 * 8;
    float s = 0.f, q = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m8; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        load8(base + i * 8, ps, np, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s += v[j];
            q += v[j] * v[j];
        }
    }
    s = block_sum(s, sh);
    q = block_sum(q, sh);
    if (threadIdx.x == 0) {
        atomicAdd(sums + 2 * smp, s);
        atomicAdd(sums + 2 * smp + 1, q);
    }
}

// y = act((x - mean_n) * rstd_n * gamma[c] + beta[c]);  relu: 0 none, 1 ReLU
__global__ void ln_apply_kernel(const bf16* __restrict__ x, long long x_ps, const float* __restrict__ sums, float inv_m,
                                float eps, const float* __restrict__ gamma, const float* __restrict__ beta, bf16* y,
                                long long y_ps, int np, long long total8, long long m8, int cg, int relu) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int smp = (int)(i / m8);
        const int ch = (int)(i % cg) * 8;
        float mean, rstd;
        ln_mean_rstd(sums, smp, inv_m, eps, mean, rstd);
        float v[8], ga[8], be[8];
        load8(x + i * 8, x_ps, np, v);
        load_f8(gamma + ch, ga);
        load_f8(beta + ch, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o = (v[j] - mean) * rstd * ga[j] + be[j];
            if (relu) o = fmaxf(o, 0.f);
            v[j] = o;
        }
        store8(y + i * 8, y_ps, np, v);
    }
}

// ---- layer_norm backward ----------------------------------------------------------------------------------------
// With g = dy * gamma, xhat = (x - mean) * rstd:  dx = rstd * (g - mean_m(g) - xhat * mean_m(g * xhat)),
// dgamma[c] = sum_{n,rows} dy * xhat, dbeta[c] = sum dy.
// Pass 1: dsums[n][2] += [sum g | sum g * xhat] of sample n, dgamma / dbeta += the per-channel sums.
// A block covers rows [r0, r0 + rows_blk) of one sample.  fixed_cg (c / 8 divides 256): a thread keeps one channel
// group, accumulates its 8 channels in registers over its rows, the row lanes are combined through shared memory
// and one atomic per channel and block goes out; otherwise (the dense layer's 16 * nf features, one row per
// sample) every item adds its channels directly.
__global__ void ln_bwd_reduce_kernel(const bf16* __restrict__ dy, long long dy_ps, const bf16* __restrict__ x,
                                     long long x_ps, const float* __restrict__ sums, float inv_m, float eps,
                                     const float* __restrict__ gamma, float* dsums, float* dgamma, float* dbeta, int np,
                                     long long rows, int cg, int rows_blk, int fixed_cg) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[256 * 8];
    __shared__ float shr[32];
    const int smp = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * rows_blk;
    const long long r1 = (r0 + rows_blk < rows) ? r0 + rows_blk : rows;
    float mean, rstd;
    ln_mean_rstd(sums, smp, inv_m, eps, mean, rstd);
    const long long base = (long long)smp * rows * cg;      // in vectors of 8
    const long long items = (r1 - r0) * cg;
    float s_g = 0.f, s_gx = 0.f;
    float a_g[8], a_b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a_g[j] = a_b[j] = 0.f;
    for (long long it = threadIdx.x; it < items; it += blockDim.x) {
        const long long vec = base + r0 * cg + it;
        const int ch = (int)(it % cg) * 8;
        float d[8], v[8], ga[8];
        load8(dy + vec * 8, dy_ps, np, d);
        load8(x + vec * 8, x_ps, np, v);
        load_f8(gamma + ch, ga);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (v[j] - mean) * rstd;
            const float g = d[j] * ga[j];
            s_g += g;
            s_gx += g * xh;
            if (fixed_cg) {
                a_g[j] += d[j] * xh;
                a_b[j] += d[j];
            } else {
                atomicAdd(dgamma + ch + j, d[j] * xh);
                atomicAdd(dbeta + ch + j, d[j]);
            }
        }
    }
    s_g = block_sum(s_g, shr);
    s_gx = block_sum(s_gx, shr);
    if (threadIdx.x == 0) {
        atomicAdd(dsums + 2 * smp, s_g);
        atomicAdd(dsums + 2 * smp + 1, s_gx);
    }
    if (fixed_cg) {
        const int lanes = blockDim.x / cg;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[threadIdx.x * 8 + j] = a_g[j];
        __syncthreads();
        if ((int)threadIdx.x < cg) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = 0.f;
                for (int l = 0; l < lanes; ++l) t += sh[(l * cg + threadIdx.x) * 8 + j];
                atomicAdd(dgamma + threadIdx.x * 8 + j, t);
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[threadIdx.x * 8 + j] = a_b[j];
        __syncthreads();
        if ((int)threadIdx.x < cg) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = 0.f;
                for (int l = 0; l < lanes; ++l) t += sh[(l * cg + threadIdx.x) * 8 + j];
                atomicAdd(dbeta + threadIdx.x * 8 + j, t);
            }
        }
    }
}

// Pass 2: dx, and dx_sum[c] += sum_{n,rows} dx (the bias gradient of the conv / dense layer in front)
__global__ void ln_bwd_apply_kernel(const bf16* __restrict__ dy, long long dy_ps, const bf16* __restrict__ x,
                                    long long x_ps, const float* __restrict__ sums, float inv_m, float eps,
                                    const float* __restrict__ gamma, const float* __restrict__ dsums, bf16* dx,
                                    long long dx_ps, float* dx_sum, int np, long long rows, int cg, int rows_blk,
                                    int fixed_cg) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[256 * 8];
    const int smp = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * rows_blk;
    const long long r1 = (r0 + rows_blk < rows) ? r0 + rows_blk : rows;
    float mean, rstd;
    ln_mean_rstd(sums, smp, inv_m, eps, mean, rstd);
    const float m_g = dsums[2 * smp] * inv_m, m_gx = dsums[2 * smp + 1] * inv_m;
    const long long base = (long long)smp * rows * cg;
    const long long items = (r1 - r0) * cg;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long it = threadIdx.x; it < items; it += blockDim.x) {
        const long long vec = base + r0 * cg + it;
        const int ch = (int)(it % cg) * 8;
        float d[8], v[8], ga[8], o[8];
        load8(dy + vec * 8, dy_ps, np, d);
        load8(x + vec * 8, x_ps, np, v);
        load_f8(gamma + ch, ga);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (v[j] - mean) * rstd;
            o[j] = rstd * (d[j] * ga[j] - m_g - xh * m_gx);
            if (dx_sum != nullptr) {
                if (fixed_cg) acc[j] += o[j];
                else atomicAdd(dx_sum + ch + j, o[j]);
            }
        }
        store8(dx + vec * 8, dx_ps, np, o);
    }
    if (dx_sum != nullptr && fixed_cg) {
        const int lanes = blockDim.x / cg;
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[threadIdx.x * 8 + j] = acc[j];
        __syncthreads();
        if ((int)threadIdx.x < cg) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = 0.f;
                for (int l = 0; l < lanes; ++l) t += sh[(l * cg + threadIdx.x) * 8 + j];
                atomicAdd(dx_sum + threadIdx.x * 8 + j, t);
            }
        }
    }
}

// ---- resampling ---------------------------------------------------------------------------------------------------
// y[n, i, j, :] = scale * x[n, i / 2, j / 2, :]   (x: h x w, y: 2h x 2w) -- resize_nearest_neighbor x2; with
// scale = 1/4 the transpose of the 2x2 average pool
// optional mask (shaped like y, post-activation values): y *= act'(mask), mask_kind 1 LeakyReLU(0.2) / 2 ReLU -- the
// pool's input-gradient and the activation derivative of the conv in front of the pool in one pass
__device__ __forceinline__ void apply_mask8(float* v, const bf16* m, long long m_ps, int np, int mask_kind) {
    float a[8];
    load8(m, m_ps, np, a);
    const float neg = (mask_kind == 1) ? 0.2f : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= (a[j] > 0.f) ? 1.f : neg;
}
__global__ void upscale2x_kernel(const bf16* __restrict__ x, long long x_ps, bf16* y, long long y_ps, int np, long long total8,
                                 int h, int w, int cg, float scale, const bf16* __restrict__ mask, long long m_ps,
                                 int mask_kind) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cg);
        long long r = i / cg;
        const int q = (int)(r % w); r /= w;
        const int p = (int)(r % h);
        const long long nn = r / h;
        float v[8];
        load8(x + i * 8, x_ps, np, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= scale;
        const long long o0 = (((nn * 2 * h + 2 * p) * 2 * w + 2 * q) * cg + c8) * 8;
        const long long rs = (long long)2 * w * cg * 8;
        const long long offs[4] = {o0, o0 + cg * 8, o0 + rs, o0 + rs + cg * 8};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = v[j];
            if (mask != nullptr) apply_mask8(u, mask + offs[t], m_ps, np, mask_kind);
            store8(y + offs[t], y_ps, np, u);
        }
    }
}
// y[n, p, q, :] = scale * sum of the 2x2 block of x   (x: h x w, y: h/2 x w/2) -- scale 1/4: tf.nn.pool AVG 2;
// scale 1: the transpose of the nearest-neighbour upscale
__global__ void pool2x_kernel(const bf16* __restrict__ x, long long x_ps, bf16* y, long long y_ps, int np, long long total8,
                              int h, int w, int cg, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    const int ho = h / 2, wo = w / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cg);
        long long r = i / cg;
        const int q = (int)(r % wo); r /= wo;
        const int p = (int)(r % ho);
        const long long nn = r / ho;
        const bf16* s = x + (((nn * h + 2 * p) * w + 2 * q) * cg + c8) * 8;
        const long long rs = (long long)w * cg * 8;
        float a[8], b[8], c[8], d[8];
        load8(s, x_ps, np, a);
        load8(s + cg * 8, x_ps, np, b);
        load8(s + rs, x_ps, np, c);
        load8(s + rs + cg * 8, x_ps, np, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = scale * ((a[j] + b[j]) + (c[j] + d[j]));
        store8(y + i * 8, y_ps, np, a);
    }
}

// out = ab[0] * x + ab[1] * z   (z == nullptr: out = ab[0] * x); ab lives in device memory so that captured graphs
// see the fade-in coefficient of the current iteration
__global__ void axpby_kernel(const bf16* __restrict__ x, long long x_ps, const bf16* __restrict__ z, long long z_ps, bf16* out,
                             long long o_ps, int np, long long n8, const float* __restrict__ ab) {
    pdl_launch_dependents();
    pdl_wait();
    const float a = ab[0], b = (z != nullptr) ? ab[1] : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float v[8], u[8];
        load8(x + i * 8, x_ps, np, v);
        if (z != nullptr) {
            load8(z + i * 8, z_ps, np, u);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = a * v[j] + b * u[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= a;
        }
        store8(out + i * 8, o_ps, np, v);
    }
}

// dst[row, d_coff + k] = src[row, s_coff + k], k < c: a channel window of one pitched buffer into another (the
// discriminator's concat buffer, pggan.py:318-322, and the gradient of its image part)
__global__ void copy_window_kernel(const bf16* __restrict__ src, long long s_ps, int s_pitch, int s_coff, bf16* dst,
                                   long long d_ps, int d_pitch, int d_coff, int np, long long total8, int cg) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cg;
        const int c8 = (int)(i % cg) * 8;
        const uint4 hi = *reinterpret_cast<const uint4*>(src + row * s_pitch + s_coff + c8);
        *reinterpret_cast<uint4*>(dst + row * d_pitch + d_coff + c8) = hi;
        if (np == 2)
            *reinterpret_cast<uint4*>(dst + d_ps + row * d_pitch + d_coff + c8) =
                *reinterpret_cast<const uint4*>(src + s_ps + row * s_pitch + s_coff + c8);
    }
}

// ---- 3-channel image ends: fp32 NHWC [.., 3] <-> planes [.., 8] (channels 3..7 zero) ---------------------------
__global__ void img_to_c8_kernel(const float* __restrict__ img, long long pixels, long long pix_per_sample,
                                 const float* __restrict__ sample_scale, bf16* dst, long long ps, int np) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        const float sc = sample_scale ? sample_scale[i / pix_per_sample] : 1.f;
        float v[8] = {img[i * 3] * sc, img[i * 3 + 1] * sc, img[i * 3 + 2] * sc, 0.f, 0.f, 0.f, 0.f, 0.f};
        store8(dst + i * 8, ps, np, v);
    }
}
__global__ void c8_to_img_kernel(const bf16* __restrict__ src, long long ps, int np, float* img, long long pixels) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        load8(src + i * 8, ps, np, v);
        img[i * 3] = v[0];
        img[i * 3 + 1] = v[1];
        img[i * 3 + 2] = v[2];
    }
}

// rows of one sample that a block of the layer-norm backward kernels covers, and whether a thread keeps one channel group
static inline void ln_geometry(long long rows, int cg, int n, int* rows_blk, int* chunks, int* fixed_cg) {
    *fixed_cg = (cg <= 256 && 256 % cg == 0) ? 1 : 0;
    const int lanes = *fixed_cg ? 256 / cg : 1;
    // enough blocks to fill the machine (~4 per SM), at least 8 row trips per block when the sample is large
    long long want = ((long long)num_sms() * 4 + n - 1) / n;
    long long rb = (rows + want - 1) / want;
    const long long min_rb = (long long)lanes * 8;
    if (rb < min_rb) rb = min_rb;
    if (rb > rows) rb = rows;
    if (rb < 1) rb = 1;
    *rows_blk = (int)rb;
    *chunks = (int)((rows + rb - 1) / rb);
}

}  // namespace t2i

using namespace t2i;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int t2i_ln_stats(const void* x, long long ps, int np, int n, long long m, float* sums, void* stream) {
    if (m % 8 != 0 || n <= 0) return fail(T2I_ERR_BAD_ARG, "ln_stats: values per sample must be a multiple of 8");
    const long long m8 = m / 8;
    long long chunks = (m8 + 256 * 8 - 1) / (256 * 8);
    const long long cap = ((long long)num_sms() * 8 + n - 1) / n;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    launch_ew(ln_stats_kernel, dim3((unsigned)chunks, (unsigned)n), dim3(256), 0, STREAM, static_cast<const bf16*>(x), ps, np, m8,
              sums);
    return check_launch("ln_stats");
}
extern "C" int t2i_ln_apply(const void* x, long long x_ps, const float* sums, float eps, const float* gamma,
                            const float* beta, void* y, long long y_ps, int np, int n, long long rows, int c, int relu,
                            void* stream) {
    if (c % 8 != 0) return fail(T2I_ERR_BAD_ARG, "ln_apply: c must be a multiple of 8");
    const long long m8 = rows * (c / 8), total8 = m8 * n;
    launch_ew(ln_apply_kernel, dim3(grid_for(total8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(x), x_ps, sums,
              1.f / (float)(rows * c), eps, gamma, beta, static_cast<bf16*>(y), y_ps, np, total8, m8, c / 8, relu);
    return check_launch("ln_apply");
}
extern "C" int t2i_ln_bwd_reduce(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* sums,
                                 float eps, const float* gamma, float* dsums, float* dgamma, float* dbeta, int np, int n,
                                 long long rows, int c, void* stream) {
    if (c % 8 != 0) return fail(T2I_ERR_BAD_ARG, "ln_bwd_reduce: c must be a multiple of 8");
    int rows_blk, chunks, fixed_cg;
    ln_geometry(rows, c / 8, n, &rows_blk, &chunks, &fixed_cg);
    launch_ew(ln_bwd_reduce_kernel, dim3((unsigned)chunks, (unsigned)n), dim3(256), 0, STREAM, static_cast<const bf16*>(dy), dy_ps,
              static_cast<const bf16*>(x), x_ps, sums, 1.f / (float)(rows * c), eps, gamma, dsums, dgamma, dbeta, np, rows,
              c / 8, rows_blk, fixed_cg);
    return check_launch("ln_bwd_reduce");
}
extern "C" int t2i_ln_bwd_apply(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* sums,
                                float eps, const float* gamma, const float* dsums, void* dx, long long dx_ps,
                                float* dx_sum, int np, int n, long long rows, int c, void* stream) {
    if (c % 8 != 0) return fail(T2I_ERR_BAD_ARG, "ln_bwd_apply: c must be a multiple of 8");
    int rows_blk, chunks, fixed_cg;
    ln_geometry(rows, c / 8, n, &rows_blk, &chunks, &fixed_cg);
    launch_ew(ln_bwd_apply_kernel, dim3((unsigned)chunks, (unsigned)n), dim3(256), 0, STREAM, static_cast<const bf16*>(dy), dy_ps,
              static_cast<const bf16*>(x), x_ps, sums, 1.f / (float)(rows * c), eps, gamma, dsums, static_cast<bf16*>(dx),
              dx_ps, dx_sum, np, rows, c / 8, rows_blk, fixed_cg);
    return check_launch("ln_bwd_apply");
}
extern "C" int t2i_upscale2x(const void* x, long long x_ps, void* y, long long y_ps, int np, int n, int h, int w, int c,
                             float scale, const void* mask, long long m_ps, int mask_kind, void* stream) {
    if (c % 8 != 0) return fail(T2I_ERR_BAD_ARG, "upscale2x: c must be a multiple of 8");
    const long long total8 = (long long)n * h * w * (c / 8);
    launch_ew(upscale2x_kernel, dim3(grid_for(total8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(x), x_ps,
              static_cast<bf16*>(y), y_ps, np, total8, h, w, c / 8, scale, static_cast<const bf16*>(mask), m_ps, mask_kind);
    return check_launch("upscale2x");
}
extern "C" int t2i_pool2x(const void* x, long long x_ps, void* y, long long y_ps, int np, int n, int h, int w, int c,
                          float scale, void* stream) {
    if (c % 8 != 0 || (h & 1) || (w & 1)) return fail(T2I_ERR_BAD_ARG, "pool2x: c must be a multiple of 8, h and w even");
    const long long total8 = (long long)n * (h / 2) * (w / 2) * (c / 8);
    launch_ew(pool2x_kernel, dim3(grid_for(total8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(x), x_ps,
              static_cast<bf16*>(y), y_ps, np, total8, h, w, c / 8, scale);
    return check_launch("pool2x");
}
extern "C" int t2i_axpby(const void* x, long long x_ps, const void* z, long long z_ps, void* out, long long o_ps, int np,
                         long long n, const float* ab, void* stream) {
    if (n % 8 != 0) return fail(T2I_ERR_BAD_ARG, "axpby: n must be a multiple of 8");
    launch_ew(axpby_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(x), x_ps,
              static_cast<const bf16*>(z), z_ps, static_cast<bf16*>(out), o_ps, np, n / 8, ab);
    return check_launch("axpby");
}
extern "C" int t2i_copy_window(const void* src, long long s_ps, int s_pitch, int s_coff, void* dst, long long d_ps,
                               int d_pitch, int d_coff, int np, long long rows, int c, void* stream) {
    if (c % 8 || s_pitch % 8 || s_coff % 8 || d_pitch % 8 || d_coff % 8)
        return fail(T2I_ERR_BAD_ARG, "copy_window: channel counts and offsets must be multiples of 8");
    if (s_coff + c > s_pitch || d_coff + c > d_pitch) return fail(T2I_ERR_BAD_ARG, "copy_window: window outside the buffer");
    const long long total8 = rows * (c / 8);
    launch_ew(copy_window_kernel, dim3(grid_for(total8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(src), s_ps, s_pitch,
              s_coff, static_cast<bf16*>(dst), d_ps, d_pitch, d_coff, np, total8, c / 8);
    return check_launch("copy_window");
}
extern "C" int t2i_img_to_c8(const float* img, int n, long long pix_per_sample, const float* sample_scale, void* dst,
                             long long ps, int np, void* stream) {
    const long long pixels = (long long)n * pix_per_sample;
    launch_ew(img_to_c8_kernel, dim3(grid_for(pixels, 256)), dim3(256), 0, STREAM, img, pixels, pix_per_sample, sample_scale,
              static_cast<bf16*>(dst), ps, np);
    return check_launch("img_to_c8");
}
extern "C" int t2i_c8_to_img(const void* src, long long ps, int np, float* img, long long pixels, void* stream) {
    launch_ew(c8_to_img_kernel, dim3(grid_for(pixels, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(src), ps, np, img,
              pixels);
    return check_launch("c8_to_img");
}
