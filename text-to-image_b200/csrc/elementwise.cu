// HBM-bound kernels of the wgancls step: coalesced 16-byte vector loads/stores, warp-shuffle and
// shared-memory reductions, fp32 arithmetic on bf16-plane storage.  Each C entry point cites the
// reference lines (relative to the reference root) whose TF ops it replaces.
#include "planes.cuh"

namespace t2i {

// ------------------------------------------------------------------------------------------
// fp32 <-> planes
__global__ void to_planes_kernel(const float* __restrict__ src, bf16* dst, long long ps, int np, long long rows,
                                 int cols, const float* __restrict__ row_scale) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n8 = rows * cols / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 8;
        const float4 a = *reinterpret_cast<const float4*>(src + e);
        const float4 b = *reinterpret_cast<const float4*>(src + e + 4);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (row_scale != nullptr) {
            const float s = row_scale[e / cols];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= s;
        }
        store8(dst + e, ps, np, v);
    }
}
__global__ void from_planes_kernel(const bf16* __restrict__ src, long long ps, int np, float* dst, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n8 = n / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        load8(src + i * 8, ps, np, v);
        *reinterpret_cast<float4*>(dst + i * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + i * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// dst[r][:] = row_scale[r] * src[r][:] (fp32; the gradient-penalty tangent seed: coef[n] * dD/dx_hat[n], model.py:62-65)
__global__ void scale_rows_kernel(const float* __restrict__ src, const float* __restrict__ row_scale, float* dst, long long rows,
                                  int cols) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n4 = rows * cols / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float sc = row_scale[i * 4 / cols];
        float4 v = *reinterpret_cast<const float4*>(src + i * 4);
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        *reinterpret_cast<float4*>(dst + i * 4) = v;
    }
}

// fp32 NHWC 3-channel image -> padded bf16 rows (planes [np][n][h][pitch], pitch = 4 * ceil((3w + 6) / 4) entries):
// B[j] = x[j - 3], zeros outside.  What the image-patch producers of conv_gemm / wgrad_gemm stream (img_patch.cuh).
__global__ void img_to_rows_kernel(const float* __restrict__ img, int n, int h, int w, const float* __restrict__ sample_scale,
                                   bf16* rows, long long ps, int np) {
    pdl_launch_dependents();
    pdl_wait();
    const int iw3 = w * 3, cpr = (iw3 + 6 + 3) / 4;             // 4-entry chunks per row
    const long long total = (long long)n * h * cpr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpr);
        const long long r = i / cpr;                              // (sample, row)
        const float sc = sample_scale ? sample_scale[r / h] : 1.f;
        const float* src = img + r * iw3;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * c - 3 + e;
            v[e] = (j >= 0 && j < iw3) ? src[j] * sc : 0.f;
        }
        bf16* dst = rows + (r * cpr + c) * 4;
        uint2 hi;
        hi.x = pack_bf16x2(v[0], v[1]);
        hi.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(dst) = hi;
        if (np == 2) {
            uint2 lo;
            lo.x = pack_bf16x2(v[0] - bf16_lo(hi.x), v[1] - bf16_hi(hi.x));
            lo.y = pack_bf16x2(v[2] - bf16_lo(hi.y), v[3] - bf16_hi(hi.y));
            *reinterpret_cast<uint2*>(dst + ps) = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------
// 3-channel 4x4/s2 patch matrix.  Row r = (n, p, q) of the (h/2 x w/2) grid, column
// (kh*4 + kw)*3 + c holds img[n, 2p-1+kh, 2q-1+kw, c] (zero outside); columns 48..63 are zero.
// A block stages the 2*PR + 2 image rows that PR rows of patches need in shared memory (coalesced 16 B loads,
// each image row read ~1.25 times instead of 2 x 12 scattered scalars per patch), then every thread assembles
// 16-byte chunks of patch rows from it.
constexpr int kIm2colPR = 4;
__global__ void im2col_k4s2_c3_kernel(const float* __restrict__ img, int n, int h, int w,
                                      const float* __restrict__ sample_scale, bf16* col, long long ps, int np) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float s_rows[];             // [2*PR + 2][w*3]
    const int hp = h / 2, wq = w / 2, rw = w * 3;
    const int groups = (hp + kIm2colPR - 1) / kIm2colPR;
    for (int blk = blockIdx.x; blk < n * groups; blk += gridDim.x) {
        const int b = blk / groups, p0 = (blk - b * groups) * kIm2colPR;
        const float sc = sample_scale ? sample_scale[b] : 1.f;
        const float* base = img + (long long)b * h * rw;
        const int ih0 = 2 * p0 - 1;
        __syncthreads();                          // the previous group's readers are done
        for (int i = threadIdx.x; i < (2 * kIm2colPR + 2) * (rw / 4); i += blockDim.x) {
            const int r = i / (rw / 4), c4 = i - r * (rw / 4);
            const int ih = ih0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ih >= 0 && ih < h) v = __ldg(reinterpret_cast<const float4*>(base + (long long)ih * rw) + c4);
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            reinterpret_cast<float4*>(s_rows + r * rw)[c4] = v;
        }
        __syncthreads();
        const int pr = min(kIm2colPR, hp - p0);
        for (int i = threadIdx.x; i < pr * wq * 8; i += blockDim.x) {
            const int chunk = i & 7, q = (i >> 3) % wq, pl = (i >> 3) / wq;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (chunk < 6) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int colj = chunk * 8 + j;
                    const int kh = colj / 12, rem = colj - kh * 12;      // rem = kw*3 + c
                    const int x0 = (2 * q - 1) * 3 + rem;               // float index inside the image row
                    if (x0 >= 0 && x0 < rw) v[j] = s_rows[(2 * pl + kh) * rw + x0];
                }
            }
            const long long r = ((long long)b * hp + p0 + pl) * wq + q;
            store8(col + r * 64 + chunk * 8, ps, np, v);
        }
    }
}
// transpose of the above: img[n, oh, ow, c] = bias[c] + sum over (kh,kw) col[(n,p,q), (kh*4+kw)*3+c]
// with oh = 2p-1+kh, ow = 2q-1+kw.
// A block produces kCol2imR image rows: it stages the R/2 + 2 rows of patches they gather from in shared memory
// (16 B loads), then each thread sums the (up to) four patches that reach one pixel.
constexpr int kCol2imR = 16;
// Wide images (256x256 in StackGAN stage-II) are cut into column strips of `ws` pixels so that the staged patches fit
// in shared memory: a strip stages sw = ws/2 + 2 patch columns starting at q0 = ow0/2 - 1 (one strip: sw = w/2, q0 = 0).
__global__ void col2im_k4s2_c3_kernel(const bf16* __restrict__ col, long long ps, int np, int n, int h, int w,
                                      const float* __restrict__ bias3, float* img, int ws, int sw) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float s_col[];               // [R/2 + 2][sw][48] fp32 values of the patch rows
    const int hp = h / 2, wq = w / 2;
    const int rgroups = (h + kCol2imR - 1) / kCol2imR;
    const int strips = (w + ws - 1) / ws;
    constexpr int PRows = kCol2imR / 2 + 2;
    const float b0 = bias3 ? bias3[0] : 0.f, b1 = bias3 ? bias3[1] : 0.f, b2 = bias3 ? bias3[2] : 0.f;
    for (int blk = blockIdx.x; blk < n * rgroups * strips; blk += gridDim.x) {
        const int strip = blk % strips, rg = (blk / strips) % rgroups, b = blk / (strips * rgroups);
        const int oh0 = rg * kCol2imR, ow0 = strip * ws;
        const int pbase = oh0 / 2 - 1;             // first patch row that can reach image row oh0
        const int qbase = strips > 1 ? ow0 / 2 - 1 : 0;
        __syncthreads();
        const int total = PRows * sw * 6;
        for (int base = 0; base < total; base += blockDim.x * 4) {      // four 16-byte loads in flight per thread
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * blockDim.x + threadIdx.x;
                const int chunk = i % 6, q = qbase + (i / 6) % sw, pl = i / (6 * sw);
                const int p = pbase + pl;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
                if (i < total && p >= 0 && p < hp && q >= 0 && q < wq)
                    load8(col + (((long long)b * hp + p) * wq + q) * 64 + chunk * 8, ps, np, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * blockDim.x + threadIdx.x;
                if (i < total) {
                    float* dst = s_col + (long long)(i / 6) * 48 + (i % 6) * 8;     // (pl * sw + q - qbase) * 48 + chunk * 8
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = v[u][j];
                }
            }
        }
        __syncthreads();
        const int rows = min(kCol2imR, h - oh0), cols = min(ws, w - ow0);
        for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
            const int ow = ow0 + i % cols, oh = oh0 + i / cols;
            float acc0 = b0, acc1 = b1, acc2 = b2;
            for (int kh = (oh + 1) & 1; kh < 4; kh += 2) {
                const int p = (oh + 1 - kh) / 2;
                if (p < 0 || p >= hp) continue;
                for (int kw = (ow + 1) & 1; kw < 4; kw += 2) {
                    const int q = (ow + 1 - kw) / 2;
                    if (q < 0 || q >= wq) continue;
                    const float* src = s_col + ((long long)(p - pbase) * sw + (q - qbase)) * 48 + (kh * 4 + kw) * 3;
                    acc0 += src[0]; acc1 += src[1]; acc2 += src[2];
                }
            }
            float* dst = img + (((long long)b * h + oh) * w + ow) * 3;
            dst[0] = acc0; dst[1] = acc1; dst[2] = acc2;
        }
    }
}

// 3x3 / stride 1 patches of a 3-channel image: row = pixel, column (kh*3 + kw)*3 + c (27 of 32 used)
__global__ void im2col_k3s1_c3_kernel(const float* __restrict__ img, int n, int h, int w, bf16* col, long long ps, int np) {
    pdl_launch_dependents();
    pdl_wait();
    const long long items = (long long)n * h * w * 4;     // (pixel, chunk of 8 columns)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i >> 2;
        const int chunk = (int)(i & 3);
        const int ow = (int)(r % w), oh = (int)((r / w) % h);
        const float* base = img + (r - (long long)oh * w - ow) * 3;     // pixel (b, 0, 0)
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int colj = chunk * 8 + j;
            const int tap = colj / 3, c = colj - tap * 3;
            const int ih = oh + tap / 3 - 1, iw = ow + tap % 3 - 1;
            v[j] = (colj < 27 && ih >= 0 && ih < h && iw >= 0 && iw < w) ? __ldg(base + ((long long)ih * w + iw) * 3 + c) : 0.f;
        }
        store8(col + r * 32 + chunk * 8, ps, np, v);
    }
}
__global__ void tanh_c3_fwd_kernel(const bf16* __restrict__ lg, long long ps, int np, float* img, long long pixels) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        load8(lg + i * 8, ps, np, v);
        img[i * 3 + 0] = tanhf(v[0]); img[i * 3 + 1] = tanhf(v[1]); img[i * 3 + 2] = tanhf(v[2]);
    }
}
__global__ void tanh_c3_bwd_kernel(const float* __restrict__ img, const float* __restrict__ dimg, bf16* dlg, long long ps,
                                   int np, long long pixels) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float y = img[i * 3 + c];
            v[c] = dimg[i * 3 + c] * (1.f - y * y);
        }
        store8(dlg + i * 8, ps, np, v);
    }
}

// ------------------------------------------------------------------------------------------
// g_net's last conv: 3 -> 3 channels, 3x3 stride 1 SAME, then tanh.  w is TF HWIO [3][3][3][3].
__global__ void conv3x3_c3_tanh_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                           const float* __restrict__ b, float* y, int n, int h, int wd) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sw[81 + 3];
    if (threadIdx.x < 81) sw[threadIdx.x] = w[threadIdx.x];
    if (threadIdx.x < 3) sw[81 + threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const long long pixels = (long long)n * h * wd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(i % wd);
        const int oh = (int)((i / wd) % h);
        const long long base = i - (long long)oh * wd - ow;  // pixel index of (b, 0, 0)
        float acc[3] = {sw[81], sw[82], sw[83]};
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = oh + kh - 1;
            if (ih < 0 || ih >= h) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = ow + kw - 1;
                if (iw < 0 || iw >= wd) continue;
                const float* xp = x + (base + (long long)ih * wd + iw) * 3;
                const float* wp = sw + (kh * 3 + kw) * 9;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float xv = xp[ci];
                    acc[0] += xv * wp[ci * 3 + 0];
                    acc[1] += xv * wp[ci * 3 + 1];
                    acc[2] += xv * wp[ci * 3 + 2];
                }
            }
        }
        y[i * 3 + 0] = tanhf(acc[0]);
        y[i * 3 + 1] = tanhf(acc[1]);
        y[i * 3 + 2] = tanhf(acc[2]);
    }
}
// backward: dl = dy * (1 - y^2); dx = conv^T(dl); dw[kh,kw,ci,co] += sum x[.+k-1, ci] dl[., co]; db += sum dl.
// Two instantiations share the staging code: WANT_DX (the input gradient, on the critical path of the backward pass:
// the 81 weights live in registers, nothing else does, several blocks per SM) and WANT_DW (the 81 weight-gradient sums
// live in registers across all tiles of the block and are reduced once at the end; runs beside the chain).  One kernel
// doing both kept 162 values per thread, ran one block per SM and read every weight from shared memory per use (60 us).
// A block walks tiles of kC9R image rows: x / dl (with a one-pixel halo) are staged in shared memory once.
constexpr int kC9R = 8;
template <bool WANT_DX, bool WANT_DW>
__global__ void __launch_bounds__(WANT_DW ? 128 : 256) conv3x3_c3_tanh_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ y, const float* __restrict__ dy, float* dx,
    float* dw, float* db, float* dx_sum, int n, int h, int wd) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float s_c9[];               // x tile [(R+2)][(wd+2)*3] (WANT_DW), dl tile the same, then w[81], red[87]
    const int tw = (wd + 2) * 3;
    float* s_x = s_c9;
    float* s_dl = s_x + (WANT_DW ? (kC9R + 2) * tw : 0);
    float* sw = s_dl + (kC9R + 2) * tw;
    float* sred = sw + 81;
    if (threadIdx.x < 81) sw[threadIdx.x] = w[threadIdx.x];
    if (threadIdx.x < 87) sred[threadIdx.x] = 0.f;
    float gw[WANT_DW ? 81 : 1];
    float wr[WANT_DX ? 81 : 1];
    float gb[3] = {0.f, 0.f, 0.f};
    float gs[3] = {0.f, 0.f, 0.f};
    if (WANT_DW) {
#pragma unroll
        for (int j = 0; j < 81; ++j) gw[j] = 0.f;
    }
    if (WANT_DX) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 81; ++j) wr[j] = sw[j];
    }
    const int groups = (h + kC9R - 1) / kC9R;
    for (int blk = blockIdx.x; blk < n * groups; blk += gridDim.x) {
        const int b = blk / groups, r0 = (blk - b * groups) * kC9R;
        const long long ibase = (long long)b * h * wd * 3;
        __syncthreads();
        // staging: 8 elements per tensor and thread with all loads issued before the first use (one memory latency per
        // trip; a one-load-at-a-time loop left the block waiting 8 latencies)
        const int total = (kC9R + 2) * tw;
        for (int base = 0; base < total; base += blockDim.x * 8) {
            float xv[8], yv[8], dv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * blockDim.x + threadIdx.x;
                const int r = i / tw, cc = i - r * tw;
                const int ih = r0 - 1 + r, iw3 = cc - 3;          // float index inside the image row
                const bool ok = i < total && ih >= 0 && ih < h && iw3 >= 0 && iw3 < wd * 3;
                const long long e = ok ? ibase + (long long)ih * wd * 3 + iw3 : 0;
                xv[u] = (WANT_DW && ok) ? __ldg(x + e) : 0.f;
                yv[u] = ok ? __ldg(y + e) : 0.f;
                dv[u] = ok ? __ldg(dy + e) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * blockDim.x + threadIdx.x;
                if (i < total) {
                    if (WANT_DW) s_x[i] = xv[u];
                    s_dl[i] = dv[u] * (1.f - yv[u] * yv[u]);
                }
            }
        }
        __syncthreads();
        const int rows = min(kC9R, h - r0);
        for (int i = threadIdx.x; i < rows * wd; i += blockDim.x) {
            const int ow = i % wd, lr = i / wd;                   // pixel (r0 + lr, ow); halo offset +1 in both directions
            const float* cx = s_x + (lr + 1) * tw + (ow + 1) * 3;
            const float* cd = s_dl + (lr + 1) * tw + (ow + 1) * 3;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (WANT_DW) {
                d0 = cd[0]; d1 = cd[1]; d2 = cd[2];               // dl of this pixel as an OUTPUT position
                gb[0] += d0; gb[1] += d1; gb[2] += d2;
            }
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int off = (kh - 1) * tw + (kw - 1) * 3;
                    if (WANT_DW) {
                        // input read by this output through tap (kh, kw): weight gradient (zero halo = padding)
                        const float x0 = cx[off], x1 = cx[off + 1], x2 = cx[off + 2];
                        float* gp = gw + (kh * 3 + kw) * 9;
                        gp[0] += x0 * d0; gp[1] += x0 * d1; gp[2] += x0 * d2;
                        gp[3] += x1 * d0; gp[4] += x1 * d1; gp[5] += x1 * d2;
                        gp[6] += x2 * d0; gp[7] += x2 * d1; gp[8] += x2 * d2;
                    }
                    if (WANT_DX) {
                        // output that reads this pixel as input through tap (kh, kw): input gradient
                        const float* wp = wr + (kh * 3 + kw) * 9;
                        const float e0 = cd[-off], e1 = cd[-off + 1], e2 = cd[-off + 2];
                        g0 += e0 * wp[0] + e1 * wp[1] + e2 * wp[2];
                        g1 += e0 * wp[3] + e1 * wp[4] + e2 * wp[5];
                        g2 += e0 * wp[6] + e1 * wp[7] + e2 * wp[8];
                    }
                }
            }
            if (WANT_DX) {
                float* dst = dx + ibase + ((long long)(r0 + lr) * wd + ow) * 3;
                dst[0] = g0; dst[1] = g1; dst[2] = g2;
                gs[0] += g0; gs[1] += g1; gs[2] += g2;
            }
        }
    }
    const int lane = threadIdx.x & 31;
    if (WANT_DW) {
#pragma unroll
        for (int j = 0; j < 81; ++j) {
            const float t = warp_sum(gw[j]);
            if (lane == 0) atomicAdd(&sred[j], t);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (WANT_DW) {
            const float t = warp_sum(gb[c]);
            if (lane == 0) atomicAdd(&sred[81 + c], t);
        }
        if (WANT_DX) {
            const float t2 = warp_sum(gs[c]);
            if (lane == 0) atomicAdd(&sred[84 + c], t2);
        }
    }
    __syncthreads();
    if (WANT_DW && threadIdx.x < 81) atomicAdd(&dw[threadIdx.x], sred[threadIdx.x]);
    else if (WANT_DW && threadIdx.x < 84) atomicAdd(&db[threadIdx.x - 81], sred[threadIdx.x]);
    else if (WANT_DX && threadIdx.x >= 84 && threadIdx.x < 87 && dx_sum != nullptr) atomicAdd(&dx_sum[threadIdx.x - 84], sred[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// Per-channel reductions over the rows of a [rows, pitch] planes tensor.  A block covers CG column
// groups (8 channels each) x RY row lanes; partial sums go through shared memory, then one fp32
// atomic per channel per block.
enum { RED_SUM = 0, RED_STATS = 1, RED_BN_BWD = 2 };

template <int MODE>
__global__ void col_reduce_kernel(const bf16* __restrict__ a, long long a_ps, const bf16* __restrict__ x, long long x_ps,
                                  const float* __restrict__ mean, const float* __restrict__ rstd, int np,
                                  long long rows, int c, int pitch, int coff, int CG, float* out0, float* out1) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sh[];  // [2][RY][CG*8]
    const int RY = blockDim.x / CG;
    const int cgl = threadIdx.x % CG, ry = threadIdx.x / CG;
    const int cg = blockIdx.x * CG + cgl;
    const int ch = cg * 8;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
    if (ch < c && ry < RY) {
        float mu[8], rs[8];
        if (MODE == RED_BN_BWD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { mu[j] = mean[ch + j]; rs[j] = rstd[ch + j]; }
        }
#pragma unroll 4
        for (long long r = (long long)blockIdx.y * RY + ry; r < rows; r += (long long)gridDim.y * RY) {
            float v[8];
            load8(a + r * pitch + coff + ch, a_ps, np, v);
            if (MODE == RED_SUM) {
#pragma unroll
                for (int j = 0; j < 8; ++j) s0[j] += v[j];
            } else if (MODE == RED_STATS) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { s0[j] += v[j]; s1[j] += v[j] * v[j]; }
            } else {
                float xv[8];
                load8(x + r * pitch + coff + ch, x_ps, np, xv);
#pragma unroll
                for (int j = 0; j < 8; ++j) { s0[j] += v[j]; s1[j] += v[j] * (xv[j] - mu[j]) * rs[j]; }
            }
        }
    }
    const int W = CG * 8;
    if (ry < RY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sh[(0 * RY + ry) * W + cgl * 8 + j] = s0[j];
            sh[(1 * RY + ry) * W + cgl * 8 + j] = s1[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) {
        const int which = i / W, col = i % W;
        if (which == 1 && MODE == RED_SUM) continue;
        const int chn = blockIdx.x * W + col;
        if (chn >= c) continue;
        float t = 0.f;
        for (int k = 0; k < RY; ++k) t += sh[(which * RY + k) * W + col];
        atomicAdd((which == 0 ? out0 : out1) + chn, t);
    }
}

template <int MODE>
static int launch_col_reduce(const void* a, long long a_ps, const void* x, long long x_ps, const float* mean,
                             const float* rstd, int np, long long rows, int c, int pitch, int coff, float* out0,
                             float* out1, cudaStream_t stream) {
    if (c % 8 || pitch % 8 || coff % 8) return fail(T2I_ERR_BAD_ARG, "col_reduce: channels must be multiples of 8");
    const int threads = 256;
    int CG = c / 8;
    if (CG > 32) CG = 32;
    // CG must divide the block: round down to a power of two
    CG = floor_pow2(CG);
    const int RY = threads / CG;
    const int gx = ceil_div(c / 8, CG);
    long long gy = (rows + RY - 1) / RY;
    const long long cap = ((long long)num_sms() * 8 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    const size_t shm = (size_t)2 * RY * CG * 8 * sizeof(float);
    launch_ew(col_reduce_kernel<MODE>, dim3(dim3(gx, (unsigned)gy)), dim3(threads), shm, stream, 
        static_cast<const bf16*>(a), a_ps, static_cast<const bf16*>(x), x_ps, mean, rstd, np, rows, c, pitch, coff, CG,
        out0, out1);
    return check_launch("col_reduce_kernel");
}

// (sum, sumsq) -> (mean, biased var, rstd), in place on the accumulators
// sums[0:c] = sum x, sums[c:2c] = sum x^2 (accumulated by col_reduce) -> mean, biased var, rstd;
// the accumulators are re-zeroed here so that the next bn_stats call needs no memset.
__global__ void bn_finalize_kernel(float* sums, float* mean, float* var, float* rstd, int c, float inv_rows, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float m = sums[i] * inv_rows;
    float v = sums[c + i] * inv_rows - m * m;
    v = fmaxf(v, 0.f);
    sums[i] = 0.f;
    sums[c + i] = 0.f;
    mean[i] = m;
    var[i] = v;
    rstd[i] = rsqrtf(v + eps);
}

// A thread owns one group of 8 channels (its BN parameters live in registers) and strides over rows;
// threadIdx.x % CG walks the channel groups so that a warp reads contiguous 16 B pieces of a row.
__global__ void bn_apply_kernel(const bf16* __restrict__ x, long long x_ps, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const bf16* __restrict__ res, long long r_ps,
                                bf16* __restrict__ y, long long y_ps, int np, long long rows, int c, int relu, int CG) {
    pdl_launch_dependents();
    pdl_wait();
    const int RY = blockDim.x / CG;
    const int ch = (blockIdx.x * CG + threadIdx.x % CG) * 8;
    const int ry = threadIdx.x / CG;
    if (ch >= c || ry >= RY) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = rstd[ch + j] * gamma[ch + j];
        sh[j] = beta[ch + j] - mean[ch + j] * sc[j];
    }
#pragma unroll 4
    for (long long r = (long long)blockIdx.y * RY + ry; r < rows; r += (long long)gridDim.y * RY) {
        const long long e = r * c + ch;
        float v[8];
        load8(x + e, x_ps, np, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = v[j] * sc[j] + sh[j];
        if (res != nullptr) {
            float rr[8];
            load8(res + e, r_ps, np, rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rr[j];
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        store8(y + e, y_ps, np, v);
    }
}

// dx = gamma * rstd * (dy - dbeta/R - xhat * dgamma/R)
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ dy, long long dy_ps, const bf16* __restrict__ x,
                                    long long x_ps, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, bf16* __restrict__ dx, long long dx_ps, int np,
                                    long long rows, int c, float inv_rows, int CG) {
    pdl_launch_dependents();
    pdl_wait();
    const int RY = blockDim.x / CG;
    const int ch = (blockIdx.x * CG + threadIdx.x % CG) * 8;
    const int ry = threadIdx.x / CG;
    if (ch >= c || ry >= RY) return;
    float mu[8], rs[8], a[8], b0[8], b1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mu[j] = mean[ch + j];
        rs[j] = rstd[ch + j];
        a[j] = gamma[ch + j] * rs[j];
        b0[j] = dbeta[ch + j] * inv_rows;
        b1[j] = dgamma[ch + j] * inv_rows;
    }
#pragma unroll 4
    for (long long r = (long long)blockIdx.y * RY + ry; r < rows; r += (long long)gridDim.y * RY) {
        const long long e = r * c + ch;
        float g[8], xv[8];
        load8(dy + e, dy_ps, np, g);
        load8(x + e, x_ps, np, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = a[j] * (g[j] - b0[j] - (xv[j] - mu[j]) * rs[j] * b1[j]);
        store8(dx + e, dx_ps, np, g);
    }
}


// Training-mode BatchNorm apply that finishes the statistics itself: sums = [sum x | sum x^2] per channel
// (accumulated by the producing conv_gemm's epilogue).  Every thread derives mean / rstd of its 8 channels;
// block row 0 also publishes mean / biased variance / rstd for the backward pass and, when asked (the G run:
// UPDATE_OPS, model.py:98,102), steps the moving statistics (utils/ops.py:20-29: decay 0.9, unbiased variance).
// Raw 16-byte register staging: the loads of a trip stay packed (4 registers per 8 values and plane) until they are
// used, which keeps these streaming kernels at 3 (np = 1) / 2 resident blocks per SM.  With the values unpacked up
// front the kernels needed 114 / 146 registers -> 2 / 1 blocks per SM and ~32 KB in flight per SM, a third of what
// the HBM latency-bandwidth product asks for (measured 1.6-2.4 TB/s).
template <int NP>
struct Raw8 {
    uint4 r[NP];
};
template <int NP>
__device__ __forceinline__ void ld_raw(const bf16* p, long long ps, Raw8<NP>& o) {
    o.r[0] = *reinterpret_cast<const uint4*>(p);
    if (NP == 2) o.r[NP - 1] = *reinterpret_cast<const uint4*>(p + ps);
}
template <int NP>
__device__ __forceinline__ void raw_f(const Raw8<NP>& o, float* v) {
    unpack8(o.r[0], v);
    if (NP == 2) {
        float l[8];
        unpack8(o.r[NP - 1], l);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += l[j];
    }
}

template <int NP, bool RES>
__global__ void __launch_bounds__(256, NP == 1 ? 3 : 2)
bn_apply_train_kernel(const bf16* __restrict__ x, long long x_ps, const float* __restrict__ sums,
                      float inv_rows, float eps, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const bf16* __restrict__ res, long long r_ps,
                      bf16* __restrict__ y, long long y_ps, long long rows, int c, int relu, int CG,
                      float* mean_out, float* rstd_out, float* var_out, float* mm, float* mv, float decay,
                      float bessel, int y_pitch, float affine_scale) {
    constexpr int U = (NP == 1 && !RES) ? 8 : 4;      // rows per thread and trip, all loads issued before the first use
    pdl_launch_dependents();
    pdl_wait();
    const int RY = blockDim.x / CG;
    const int ch = (blockIdx.x * CG + threadIdx.x % CG) * 8;
    const int ry = threadIdx.x / CG;
    if (ch >= c || ry >= RY) return;
    const float act_neg = (relu == 2) ? 0.2f : 0.f;      // relu: 0 none, 1 ReLU, 2 LeakyReLU(0.2)
    float sc[8], sh[8];
    {
        float s1[8], s2[8], ga[8], be[8];
        load_f8(sums + ch, s1);
        load_f8(sums + c + ch, s2);
        load_f8(gamma + ch, ga);
        load_f8(beta + ch, be);
        const bool publish = blockIdx.y == 0 && ry == 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float m = s1[j] * inv_rows;
            const float v = fmaxf(s2[j] * inv_rows - m * m, 0.f);
            const float rs = rsqrtf(v + eps);
            if (publish) {
                mean_out[ch + j] = m;
                var_out[ch + j] = v;
                rstd_out[ch + j] = rs;
                if (mm != nullptr) {
                    mm[ch + j] = decay * mm[ch + j] + (1.f - decay) * m;
                    mv[ch + j] = decay * mv[ch + j] + (1.f - decay) * v * bessel;
                }
            }
            sc[j] = rs * ga[j] * affine_scale;
            sh[j] = be[j] * affine_scale - m * sc[j];
        }
    }
    const long long stride = (long long)gridDim.y * RY;
    for (long long r0 = (long long)blockIdx.y * RY + ry; r0 < rows; r0 += U * stride) {
        Raw8<NP> xr[U], rr[RES ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * stride;
            if (r < rows) {
                ld_raw<NP>(x + r * c + ch, x_ps, xr[u]);
                if (RES) ld_raw<NP>(res + r * c + ch, r_ps, rr[RES ? u : 0]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * stride;
            if (r < rows) {
                float v[8];
                raw_f<NP>(xr[u], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = v[j] * sc[j] + sh[j];
                if (RES) {
                    float q[8];
                    raw_f<NP>(rr[RES ? u : 0], q);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += q[j];
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], act_neg * v[j]);
                }
                store8(y + r * y_pitch + ch, y_ps, NP, v);
            }
        }
    }
}

// BatchNorm input-gradient with the reductions already done by the kernel that produced dy:
// dbeta = sum dy, dot = sum dy * x (raw pre-normalisation input).  dgamma = rstd * (dot - mean * dbeta);
// dx = gamma * rstd * (dy - dbeta/R - xhat * dgamma/R) = A * dy + B * x + C per channel.  Block row 0 adds dgamma into
// the gradient buffer; dx_sum (optional) accumulates the per-channel sum of dx = the bias gradient of the conv in
// front of the BN.
template <int NP>
__global__ void __launch_bounds__(256, NP == 1 ? 3 : 2)
bn_bwd_fused_kernel(const bf16* __restrict__ dy, long long dy_ps, const bf16* __restrict__ x,
                    long long x_ps, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ dot,
                    const float* __restrict__ dbeta, float* dgamma_out, float* dbeta_out, float out_scale,
                    int dot_normalised, bf16* __restrict__ dx, long long dx_ps, float* dx_sum,
                    long long rows, int c, float inv_rows, int CG, int dy_pitch, float affine_scale) {
    constexpr int U = 4;
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[256 * 8];
    const int RY = blockDim.x / CG;
    const int cgl = threadIdx.x % CG;
    const int ch = (blockIdx.x * CG + cgl) * 8;
    const int ry = threadIdx.x / CG;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ch < c && ry < RY) {
        float A[8], B[8], C[8];
        {
            float mu[8], rs[8], ga[8], dt[8], db8[8];
            load_f8(mean + ch, mu);
            load_f8(rstd + ch, rs);
            load_f8(gamma + ch, ga);
            load_f8(dot + ch, dt);
            load_f8(dbeta + ch, db8);
            const bool publish = blockIdx.y == 0 && ry == 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float db = db8[j];
                const float dg = dot_normalised ? dt[j] : rs[j] * (dt[j] - mu[j] * db);
                if (publish) {
                    dgamma_out[ch + j] += dg * out_scale;
                    if (dbeta_out != nullptr) dbeta_out[ch + j] += db * out_scale;
                }
                const float a = ga[j] * affine_scale * rs[j];
                A[j] = a;
                B[j] = -a * rs[j] * dg * inv_rows;
                C[j] = -a * db * inv_rows - B[j] * mu[j];
            }
        }
        const long long stride = (long long)gridDim.y * RY;
        for (long long r0 = (long long)blockIdx.y * RY + ry; r0 < rows; r0 += U * stride) {
            Raw8<NP> gr[U], xr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * stride;
                if (r < rows) {
                    ld_raw<NP>(dy + r * dy_pitch + ch, dy_ps, gr[u]);
                    ld_raw<NP>(x + r * c + ch, x_ps, xr[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * stride;
                if (r < rows) {
                    float g[8], xv[8];
                    raw_f<NP>(gr[u], g);
                    raw_f<NP>(xr[u], xv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        g[j] = A[j] * g[j] + B[j] * xv[j] + C[j];
                        acc[j] += g[j];
                    }
                    store8(dx + r * c + ch, dx_ps, NP, g);
                }
            }
        }
    }
    if (dx_sum == nullptr) return;
    const int W = CG * 8;
    if (ry < RY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[ry * W + cgl * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < W; col += blockDim.x) {
        const int chn = blockIdx.x * W + col;
        if (chn >= c) continue;
        float t = 0.f;
        for (int k = 0; k < RY; ++k) t += sh[k * W + col];
        atomicAdd(dx_sum + chn, t);
    }
}

// launch geometry shared by the row-striding kernels above
static inline void rowwise_geometry(long long rows, int c, int threads, int* CG, dim3* grid, int rows_per_thread = 1) {
    int cg = c / 8;
    int g = cg > 32 ? 32 : floor_pow2(cg);
    const int RY = threads / g;
    const int gx = ceil_div(cg, g);
    // rows_per_thread > 1: the per-thread parameter prologue (8 channels x several arrays) is amortised over that
    // many 16-byte payload accesses instead of being paid once per access
    long long gy = (rows + (long long)RY * rows_per_thread - 1) / ((long long)RY * rows_per_thread);
    const long long cap = ((long long)num_sms() * 16 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    *CG = g;
    *grid = dim3(gx, (unsigned)gy);
}

__global__ void bn_update_moving_kernel(float* mm, float* mv, const float* mean, const float* var, int c, float decay,
                                        float bessel) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    mm[i] = decay * mm[i] + (1.f - decay) * mean[i];
    mv[i] = decay * mv[i] + (1.f - decay) * var[i] * bessel;
}

__global__ void act_bwd_kernel(const bf16* __restrict__ dy, long long dy_ps, const bf16* __restrict__ y, long long y_ps,
                               bf16* dst, long long dst_ps, int np, long long n8, float neg) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float g[8], a[8];
        load8(dy + i * 8, dy_ps, np, g);
        load8(y + i * 8, y_ps, np, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= (a[j] > 0.f) ? 1.f : neg;
        store8(dst + i * 8, dst_ps, np, g);
    }
}

// ------------------------------------------------------------------------------------------
// d_net embedding replicate / reduce
__global__ void embed_tile_kernel(const bf16* __restrict__ e, long long e_ps, bf16* cat, long long cat_ps, int np, int s,
                                  int c, int pitch, int coff, int hw) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg = c / 8;
    const long long items = (long long)s * hw * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        const long long pix = i / cg;  // sample * hw + pos
        const long long smp = pix / hw;
        float v[8];
        load8(e + smp * c + g * 8, e_ps, np, v);
        store8(cat + pix * pitch + coff + g * 8, cat_ps, np, v);
    }
}
__global__ void embed_reduce_kernel(const bf16* __restrict__ dcat, long long dcat_ps, bf16* de, long long de_ps, int np,
                                    int s, int c, int pitch, int coff, int hw) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg = c / 8;
    const long long items = (long long)s * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        const long long smp = i / cg;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int pos = 0; pos < hw; ++pos) {
            float v[8];
            load8(dcat + (smp * hw + pos) * pitch + coff + g * 8, dcat_ps, np, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
        store8(de + smp * c + g * 8, de_ps, np, acc);
    }
}

// ------------------------------------------------------------------------------------------
// d_net output layer: per-sample dot product over k = 4*4*C values (NHWC order == TF HWIO order)
__global__ void dout_fwd_kernel(const bf16* __restrict__ a, long long a_ps, int np, const float* __restrict__ w,
                                const float* __restrict__ b, float* logit, int k) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    const long long s = blockIdx.x;
    float acc = 0.f;
    for (int i = threadIdx.x * 8; i < k; i += blockDim.x * 8) {
        float v[8];
        load8(a + s * k + i, a_ps, np, v);
        const float4 w0 = *reinterpret_cast<const float4*>(w + i);
        const float4 w1 = *reinterpret_cast<const float4*>(w + i + 4);
        acc += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
    }
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) logit[s] = t + b[0];
}
__global__ void dout_bwd_data_kernel(const bf16* __restrict__ a, long long a_ps, int np, const float* __restrict__ w,
                                     const float* __restrict__ seed, bf16* da, long long da_ps, int s, int k) {
    pdl_launch_dependents();
    pdl_wait();
    const int kg = k / 8;
    const long long items = (long long)s * kg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % kg);
        const long long smp = i / kg;
        const float sd = seed[smp];
        float v[8];
        load8(a + i * 8, a_ps, np, v);
        const float4 w0 = *reinterpret_cast<const float4*>(w + g * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(w + g * 8 + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = sd * ww[j] * ((v[j] > 0.f) ? 1.f : 0.2f);
        store8(da + i * 8, da_ps, np, v);
    }
}
// dw[k] += sum_s seed[s] * a[s, k]; blockIdx.y strides over samples
__global__ void dout_bwd_weight_kernel(const bf16* __restrict__ a, long long a_ps, int np, const float* __restrict__ seed,
                                       float* dw, int s, int k) {
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * 8 >= k) return;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int smp = blockIdx.y; smp < s; smp += gridDim.y) {
        float v[8];
        load8(a + (long long)smp * k + g * 8, a_ps, np, v);
        const float sd = seed[smp];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += sd * v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dw + g * 8 + j, acc[j]);
}
__global__ void seed_sum_kernel(const float* __restrict__ seed, int n, float* out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += seed[i];
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(out, t);
}

// ------------------------------------------------------------------------------------------
// gradient penalty path
__global__ void gp_interp_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ eps,
                                 float* xhat, long long n4, int per_sample4) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float e = eps[i / per_sample4];
        const float4 a = reinterpret_cast<const float4*>(g)[i];
        const float4 b = reinterpret_cast<const float4*>(x)[i];
        float4 o;
        o.x = e * a.x + (1.f - e) * b.x; o.y = e * a.y + (1.f - e) * b.y;
        o.z = e * a.z + (1.f - e) * b.z; o.w = e * a.w + (1.f - e) * b.w;
        reinterpret_cast<float4*>(xhat)[i] = o;
    }
}
__global__ void gp_penalty_kernel(const float* __restrict__ grad, int per_sample, float weight, float inv_batch,
                                  float* slope, float* coef, float* pen_sum) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    const long long b = blockIdx.x;
    const float* gp = grad + b * per_sample;
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < per_sample; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(gp + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        const float s = sqrtf(t);
        const float ex = fmaxf(s - 1.f, 0.f);
        slope[b] = s;
        coef[b] = (ex > 0.f) ? weight * 2.f * ex / s * inv_batch : 0.f;
        atomicAdd(pen_sum, ex * ex);
    }
}

// ------------------------------------------------------------------------------------------
// conditioning augmentation
// ms = [mean | log_sigma] is an fp32 tensor (t2i_dense_f32): log_sigma feeds exp(), see img_gemm.cu
__global__ void ca_fwd_kernel(const float* __restrict__ ms, const float* __restrict__ z,
                              const float* __restrict__ tn, bf16* zc, long long zc_ps, int np, int b, int z_dim, int ce,
                              float* kl_sum) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    const int width = z_dim + ce;
    float kl = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)b * width;
         i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % width);
        const long long row = i / width;
        float v;
        if (col < z_dim) {
            v = z[row * z_dim + col];
        } else {
            const int j = col - z_dim;
            const float mean = ms[row * 2 * ce + j];
            const float ls = ms[row * 2 * ce + ce + j];
            v = mean + expf(ls) * tn[row * ce + j];
            kl += -ls + 0.5f * (-1.f + expf(2.f * ls) + mean * mean);
        }
        store1(zc + i, zc_ps, np, v);
    }
    const float t = block_sum(kl, sh);
    if (threadIdx.x == 0 && kl_sum != nullptr) atomicAdd(kl_sum, t);
}
__global__ void ca_bwd_kernel(const float* __restrict__ ms, const bf16* __restrict__ dzc, long long dzc_ps,
                              const float* __restrict__ tn, bf16* dms, long long dms_ps, int np, int b, int z_dim, int ce,
                              float kl_scale) {
    pdl_launch_dependents();
    pdl_wait();
    const int width = z_dim + ce;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)b * ce;
         i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % ce);
        const long long row = i / ce;
        const float mean = ms[row * 2 * ce + j];
        const float ls = ms[row * 2 * ce + ce + j];
        const float dc = load1(dzc + row * width + z_dim + j, dzc_ps, np);
        float dmean = dc + kl_scale * mean;
        float dls = dc * tn[row * ce + j] * expf(ls) + kl_scale * (expf(2.f * ls) - 1.f);
        dmean *= (mean > 0.f) ? 1.f : 0.2f;  // LeakyReLU on both heads (model.py:113-114)
        dls *= (ls > 0.f) ? 1.f : 0.2f;
        store1(dms + row * 2 * ce + j, dms_ps, np, dmean);
        store1(dms + row * 2 * ce + ce + j, dms_ps, np, dls);
    }
}

// ------------------------------------------------------------------------------------------
// scalars
__global__ void d_seeds_kernel(const float* kt, float* seed, int b, float inv_gb) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * b) return;
    const float k = kt[0];
    const int seg = i / b;
    // D_loss = -(real - fake) - kt (real - mismatch) + gp terms   (model.py:82-91)
    seed[i] = (seg == 0) ? inv_gb : (seg == 1) ? -(1.f + k) * inv_gb : (seg == 2) ? k * inv_gb : 1.f;
}
__global__ void d_sums_kernel(const float* __restrict__ logit, int b, float* sums) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < b; i += blockDim.x) {
        acc[0] += logit[i];
        acc[1] += logit[b + i];
        const float m = logit[2 * b + i];
        acc[2] += m;
        acc[3] += m * m;
    }
    for (int j = 0; j < 4; ++j) {
        const float t = block_sum(acc[j], sh);
        if (threadIdx.x == 0) atomicAdd(&sums[j], t);
    }
}
__global__ void d_scalars_kernel(const float* sums, float* kt, float* sc, float inv_gb, float gp_weight, float kt_lr) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float fake = sums[0] * inv_gb, real = sums[1] * inv_gb, mis = sums[2] * inv_gb;
    const float reg = sums[3] * inv_gb, gp = sums[4] * inv_gb, gp2 = sums[5] * inv_gb;
    const float k = kt[0];
    const float wdist = real - fake, wdist2 = real - mis;
    const float bal = k * wdist2 - wdist;
    sc[T2I_S_D_LOSS_REAL] = real;
    sc[T2I_S_D_LOSS_FAKE] = fake;
    sc[T2I_S_D_LOSS_MISMATCH] = mis;
    sc[T2I_S_WDIST] = wdist;
    sc[T2I_S_WDIST2] = wdist2;
    sc[T2I_S_REG_LOSS] = reg;
    sc[T2I_S_BALANCE_LOSS] = bal * bal;
    sc[T2I_S_REAL_GP] = gp;
    sc[T2I_S_REAL_GP2] = gp2;
    sc[T2I_S_D_LOSS] = -wdist - k * wdist2 + gp_weight * (gp + gp2);
    const float kt_grad = 2.f * bal * wdist2;   // d balance_loss / d kt   (model.py:85,100)
    sc[T2I_S_KT_GRAD] = kt_grad;
    const float nk = k - kt_lr * kt_grad;
    kt[0] = nk;
    sc[T2I_S_KT] = nk;
}
// StackGAN stage-I losses (models/stackgan/stageI/trainer.py:21-44): per-sample backward seed of
// weight * mean_b sigmoid_cross_entropy(logit, label) and the running sum of the cross-entropy terms.
__global__ void ce_seeds_kernel(const float* __restrict__ logit, int n, float label, float weight, float inv_gb, float* seed,
                                float* loss_sum) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = logit[i];
        const float sg = 1.f / (1.f + expf(-x));
        seed[i] = weight * (sg - label) * inv_gb;
        acc += fmaxf(x, 0.f) - x * label + log1pf(expf(-fabsf(x)));
    }
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0 && loss_sum != nullptr) atomicAdd(loss_sum, t);
}
// sums: [0] CE(synthetic, 0)  [1] CE(real match, 0.9)  [2] CE(real mismatch, 0)  [3] CE(synthetic, 1)  [4] KL terms
__global__ void s1_scalars_kernel(const float* sums, float* sc, float inv_gb, float inv_gb_ce, float alpha, float kl_coeff,
                                  int which) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (which == 0) {
        const float syn = sums[0] * inv_gb, real = sums[1] * inv_gb, mis = sums[2] * inv_gb;
        sc[0] = real + alpha * mis + (1.f - alpha) * syn;      // D_loss
        sc[1] = syn; sc[2] = real; sc[3] = mis;
    } else {
        const float gan = sums[3] * inv_gb, kl = sums[4] * inv_gb_ce;
        sc[4] = gan + kl_coeff * kl;                           // G_loss
        sc[5] = gan; sc[6] = kl;
    }
}
__global__ void g_sums_kernel(const float* __restrict__ logit, int b, float* sums) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < b; i += blockDim.x) acc += logit[i];
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(&sums[0], t);
}
__global__ void g_scalars_kernel(const float* sums, float* sc, float inv_gb, float inv_gb_ce, float kl_coeff) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float fake = sums[0] * inv_gb;
    const float kl = sums[1] * inv_gb_ce;
    sc[T2I_S_G_KL_LOSS] = kl;
    sc[T2I_S_G_LOSS] = -fake + kl_coeff * kl;   // model.py:92
}

// ------------------------------------------------------------------------------------------
// weights: fp32 [taps][cout][cin] -> planes fwd (same layout) and bwd ([taps][cin][cout]); 32x32 tiles
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, bf16* fwd, long long fwd_ps, bf16* bwd,
                                   long long bwd_ps, int np) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float tile[32][33];
    const int tap = blockIdx.z;
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    const float* src = w + (long long)tap * cout * cin;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int co = co0 + r, ci = ci0 + threadIdx.x;
        float v = 0.f;
        if (co < cout && ci < cin) {
            v = src[(long long)co * cin + ci];
            if (fwd != nullptr) store1(fwd + ((long long)tap * cout + co) * cin + ci, fwd_ps, np, v);
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    if (bwd != nullptr) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int ci = ci0 + r, co = co0 + threadIdx.x;
            if (co < cout && ci < cin) store1(bwd + ((long long)tap * cin + ci) * cout + co, bwd_ps, np, tile[threadIdx.x][r]);
        }
    }
}

__global__ void adam_tf_kernel(float* theta, const float* __restrict__ grad, float* m, float* v, long long n,
                               const float* __restrict__ lr_t_dev, float b1, float b2, float eps, float gs, bf16* packed,
                               long long packed_ps, int np) {
    pdl_launch_dependents();
    pdl_wait();
    const float lr_t = lr_t_dev[0];
    const long long n8 = n / 8;
    const bool use_m = (b1 != 0.f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float g[8], mm[8], vv[8], th[8];
        *reinterpret_cast<float4*>(g) = reinterpret_cast<const float4*>(grad)[2 * i];
        *reinterpret_cast<float4*>(g + 4) = reinterpret_cast<const float4*>(grad)[2 * i + 1];
        *reinterpret_cast<float4*>(vv) = reinterpret_cast<float4*>(v)[2 * i];
        *reinterpret_cast<float4*>(vv + 4) = reinterpret_cast<float4*>(v)[2 * i + 1];
        *reinterpret_cast<float4*>(th) = reinterpret_cast<float4*>(theta)[2 * i];
        *reinterpret_cast<float4*>(th + 4) = reinterpret_cast<float4*>(theta)[2 * i + 1];
        if (use_m) {
            *reinterpret_cast<float4*>(mm) = reinterpret_cast<float4*>(m)[2 * i];
            *reinterpret_cast<float4*>(mm + 4) = reinterpret_cast<float4*>(m)[2 * i + 1];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gj = g[j] * gs;
            const float mj = use_m ? b1 * mm[j] + (1.f - b1) * gj : gj;
            mm[j] = mj;
            vv[j] = b2 * vv[j] + (1.f - b2) * gj * gj;
            th[j] -= lr_t * mj / (sqrtf(vv[j]) + eps);
        }
        if (use_m) {
            reinterpret_cast<float4*>(m)[2 * i] = *reinterpret_cast<float4*>(mm);
            reinterpret_cast<float4*>(m)[2 * i + 1] = *reinterpret_cast<float4*>(mm + 4);
        }
        reinterpret_cast<float4*>(v)[2 * i] = *reinterpret_cast<float4*>(vv);
        reinterpret_cast<float4*>(v)[2 * i + 1] = *reinterpret_cast<float4*>(vv + 4);
        reinterpret_cast<float4*>(theta)[2 * i] = *reinterpret_cast<float4*>(th);
        reinterpret_cast<float4*>(theta)[2 * i + 1] = *reinterpret_cast<float4*>(th + 4);
        if (packed != nullptr) store8(packed + i * 8, packed_ps, np, th);
    }
    // tail (n not a multiple of 8)
    for (long long i = n8 * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gj = grad[i] * gs;
        const float mj = use_m ? b1 * m[i] + (1.f - b1) * gj : gj;
        if (use_m) m[i] = mj;
        v[i] = b2 * v[i] + (1.f - b2) * gj * gj;
        theta[i] -= lr_t * mj / (sqrtf(v[i]) + eps);
        if (packed != nullptr) store1(packed + i, packed_ps, np, theta[i]);
    }
}

}  // namespace t2i

using namespace t2i;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int t2i_to_planes(const float* src, void* dst, long long ps, int np, long long rows, int cols,
                             const float* row_scale, void* stream) {
    if ((rows * cols) % 8 != 0 || cols % 8 != 0) return fail(T2I_ERR_BAD_ARG, "to_planes: cols must be a multiple of 8");
    launch_ew(to_planes_kernel, dim3(grid_for(rows * cols / 8, 256)), dim3(256), 0, STREAM, src, static_cast<bf16*>(dst), ps, np, rows, cols, row_scale);
    return check_launch("to_planes");
}
extern "C" int t2i_from_planes(const void* src, long long ps, int np, float* dst, long long n, void* stream) {
    if (n % 8 != 0) return fail(T2I_ERR_BAD_ARG, "from_planes: n must be a multiple of 8");
    launch_ew(from_planes_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(src), ps, np, dst, n);
    return check_launch("from_planes");
}
extern "C" int t2i_scale_rows(const float* src, const float* row_scale, float* dst, long long rows, int cols, void* stream) {
    if (cols % 4 != 0) return fail(T2I_ERR_BAD_ARG, "scale_rows: cols=%d must be a multiple of 4", cols);
    launch_ew(scale_rows_kernel, dim3(grid_for(rows * cols / 4, 256)), dim3(256), 0, STREAM, src, row_scale, dst, rows, cols);
    return check_launch("scale_rows");
}
extern "C" int t2i_img_to_rows(const float* img, int n, int h, int w, const float* sample_scale, void* rows,
                               long long plane_stride, int np, void* stream) {
    const long long chunks = (long long)n * h * ((w * 3 + 6 + 3) / 4);
    launch_ew(img_to_rows_kernel, dim3(grid_for(chunks, 256)), dim3(256), 0, STREAM, img, n, h, w, sample_scale,
              static_cast<bf16*>(rows), plane_stride, np);
    return check_launch("img_to_rows");
}
extern "C" int t2i_im2col_k4s2_c3(const float* img, int n, int h, int w, const float* sample_scale, void* col,
                                  long long ps, int np, void* stream) {
    if ((h & 1) || (w & 1)) return fail(T2I_ERR_BAD_ARG, "im2col: odd extent");
    if (w % 4) return fail(T2I_ERR_BAD_ARG, "im2col: w must be a multiple of 4");
    const size_t shm = (size_t)(2 * kIm2colPR + 2) * w * 3 * sizeof(float);
    if (shm > 48 * 1024) return fail(T2I_ERR_BAD_ARG, "im2col: image rows too wide (%d)", w);
    const long long groups = (long long)n * ceil_div(h / 2, kIm2colPR);
    const long long cap = (long long)num_sms() * 8;
    launch_ew(im2col_k4s2_c3_kernel, dim3((unsigned)(groups < cap ? groups : cap)), dim3(256), shm, STREAM, img, n, h, w, sample_scale,
                                                                                      static_cast<bf16*>(col), ps, np);
    return check_launch("im2col_k4s2_c3");
}
extern "C" int t2i_col2im_k4s2_c3(const void* col, long long ps, int np, int n, int h, int w, const float* bias3,
                                  float* img, void* stream) {
    if ((h & 1) || (w & 1)) return fail(T2I_ERR_BAD_ARG, "col2im: odd extent");
    // one strip when the whole patch row fits (w <= 64), otherwise 64-pixel column strips with a one-patch halo
    const int ws = w <= 64 ? w : 64;
    const int sw = w <= 64 ? w / 2 : ws / 2 + 2;
    const size_t shm = (size_t)(kCol2imR / 2 + 2) * sw * 48 * sizeof(float);
    if (shm > 96 * 1024) return fail(T2I_ERR_BAD_ARG, "col2im: image rows too wide (%d)", w);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(col2im_k4s2_c3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_done = true;
    }
    const long long groups = (long long)n * ceil_div(h, kCol2imR) * ceil_div(w, ws);
    const long long cap = (long long)num_sms() * 4;
    launch_ew(col2im_k4s2_c3_kernel, dim3((unsigned)(groups < cap ? groups : cap)), dim3(256), shm, STREAM, static_cast<const bf16*>(col), ps, np,
                                                                                      n, h, w, bias3, img, ws, sw);
    return check_launch("col2im_k4s2_c3");
}
extern "C" int t2i_im2col_k3s1_c3(const float* img, int n, int h, int w, void* col, long long ps, int np, void* stream) {
    launch_ew(im2col_k3s1_c3_kernel, dim3(grid_for((long long)n * h * w * 4, 256, 16)), dim3(256), 0, STREAM, img, n, h, w,
              static_cast<bf16*>(col), ps, np);
    return check_launch("im2col_k3s1_c3");
}
extern "C" int t2i_tanh_c3_fwd(const void* logits8, long long ps, int np, float* img, long long pixels, void* stream) {
    launch_ew(tanh_c3_fwd_kernel, dim3(grid_for(pixels, 256, 16)), dim3(256), 0, STREAM, static_cast<const bf16*>(logits8), ps,
              np, img, pixels);
    return check_launch("tanh_c3_fwd");
}
extern "C" int t2i_tanh_c3_bwd(const float* img, const float* dimg, void* dlogits8, long long ps, int np, long long pixels,
                               void* stream) {
    launch_ew(tanh_c3_bwd_kernel, dim3(grid_for(pixels, 256, 16)), dim3(256), 0, STREAM, img, dimg, static_cast<bf16*>(dlogits8),
              ps, np, pixels);
    return check_launch("tanh_c3_bwd");
}
extern "C" int t2i_conv3x3_c3_tanh_fwd(const float* x, const float* w, const float* b, float* y, int n, int h, int wd,
                                       void* stream) {
    launch_ew(conv3x3_c3_tanh_fwd_kernel, dim3(grid_for((long long)n * h * wd, 256, 16)), dim3(256), 0, STREAM, x, w, b, y, n, h, wd);
    return check_launch("conv3x3_c3_tanh_fwd");
}
extern "C" int t2i_conv3x3_c3_tanh_bwd(const float* x, const float* w, const float* y, const float* dy, float* dx,
                                       float* dw, float* db, float* dx_sum, int n, int h, int wd, void* stream) {
    // dx != NULL: the input gradient (+ dx_sum); dw != NULL (with db): the weight / bias gradients; both: two launches
    if (dx == nullptr && dw == nullptr) return fail(T2I_ERR_BAD_ARG, "conv3x3_c3_tanh_bwd: neither dx nor dw requested");
    if (dw != nullptr && (db == nullptr || x == nullptr)) return fail(T2I_ERR_BAD_ARG, "conv3x3_c3_tanh_bwd: dw needs db and x");
    const size_t tile = (size_t)(kC9R + 2) * (wd + 2) * 3;
    if ((2 * tile + 81 + 87) * sizeof(float) > 48 * 1024) return fail(T2I_ERR_BAD_ARG, "conv3x3_c3_tanh_bwd: image rows too wide (%d)", wd);
    const long long groups = (long long)n * ceil_div(h, kC9R);
    if (dx != nullptr) {
        const long long cap = 4LL * num_sms();
        launch_ew(conv3x3_c3_tanh_bwd_kernel<true, false>, dim3((unsigned)(groups < cap ? groups : cap)), dim3(256),
                  (tile + 81 + 87) * sizeof(float), STREAM, x, w, y, dy, dx, dw, db, dx_sum, n, h, wd);
        int rc = check_launch("conv3x3_c3_tanh_bwd (dx)");
        if (rc != T2I_OK) return rc;
    }
    if (dw != nullptr) {
        const long long cap = 4LL * num_sms();      // 127 registers x 128 threads: four blocks per SM
        launch_ew(conv3x3_c3_tanh_bwd_kernel<false, true>, dim3((unsigned)(groups < cap ? groups : cap)), dim3(128),
                  (2 * tile + 81 + 87) * sizeof(float), STREAM, x, w, y, dy, dx, dw, db, dx_sum, n, h, wd);
        return check_launch("conv3x3_c3_tanh_bwd (dw)");
    }
    return T2I_OK;
}
extern "C" int t2i_colsum(const void* src, long long ps, int np, long long rows, int c, int pitch, int coff, float* out,
                          void* stream) {
    return launch_col_reduce<RED_SUM>(src, ps, nullptr, 0, nullptr, nullptr, np, rows, c, pitch, coff, out, nullptr, STREAM);
}
extern "C" int t2i_bn_stats(const void* x, long long ps, int np, long long rows, int c, float* sums, float* mean,
                            float* rstd, float* var, float eps, void* stream) {
    int rc = launch_col_reduce<RED_STATS>(x, ps, nullptr, 0, nullptr, nullptr, np, rows, c, c, 0, sums, sums + c, STREAM);
    if (rc != T2I_OK) return rc;
    launch_ew(bn_finalize_kernel, dim3(ceil_div(c, 256)), dim3(256), 0, STREAM, sums, mean, var, rstd, c, 1.f / (float)rows, eps);
    return check_launch("bn_finalize");
}
extern "C" int t2i_bn_apply(const void* x, long long x_ps, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const void* residual, long long r_ps, void* y, long long y_ps, int np,
                            long long rows, int c, int relu, void* stream) {
    if (c % 8) return fail(T2I_ERR_BAD_ARG, "bn_apply: c must be a multiple of 8");
    int CG;
    dim3 grid;
    rowwise_geometry(rows, c, 256, &CG, &grid);
    launch_ew(bn_apply_kernel, dim3(grid), dim3(256), 0, STREAM, 
        static_cast<const bf16*>(x), x_ps, mean, rstd, gamma, beta, static_cast<const bf16*>(residual), r_ps,
        static_cast<bf16*>(y), y_ps, np, rows, c, relu, CG);
    return check_launch("bn_apply");
}
extern "C" int t2i_bn_bwd_reduce(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* mean,
                                 const float* rstd, int np, long long rows, int c, float* dgamma, float* dbeta,
                                 void* stream) {
    return launch_col_reduce<RED_BN_BWD>(dy, dy_ps, x, x_ps, mean, rstd, np, rows, c, c, 0, dbeta, dgamma, STREAM);
}
extern "C" int t2i_bn_bwd_apply(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* mean,
                                const float* rstd, const float* gamma, const float* dgamma, const float* dbeta, void* dx,
                                long long dx_ps, int np, long long rows, int c, void* stream) {
    if (c % 8) return fail(T2I_ERR_BAD_ARG, "bn_bwd_apply: c must be a multiple of 8");
    int CG;
    dim3 grid;
    rowwise_geometry(rows, c, 256, &CG, &grid);
    launch_ew(bn_bwd_apply_kernel, dim3(grid), dim3(256), 0, STREAM, 
        static_cast<const bf16*>(dy), dy_ps, static_cast<const bf16*>(x), x_ps, mean, rstd, gamma, dgamma, dbeta,
        static_cast<bf16*>(dx), dx_ps, np, rows, c, 1.f / (float)rows, CG);
    return check_launch("bn_bwd_apply");
}
extern "C" int t2i_bn_apply_train(const void* x, long long x_ps, const float* sums, float eps, const float* gamma,
                                  const float* beta, const void* residual, long long r_ps, void* y, long long y_ps, int np,
                                  long long rows, int c, int relu, float* mean, float* rstd, float* var,
                                  float* moving_mean, float* moving_var, float decay, long long stat_rows, int y_pitch,
                                  float affine_scale, void* stream) {
    if (c % 8 || y_pitch % 8) return fail(T2I_ERR_BAD_ARG, "bn_apply_train: c and y_pitch must be multiples of 8");
    if (y_pitch == 0) y_pitch = c;
    if (y_pitch < c) return fail(T2I_ERR_BAD_ARG, "bn_apply_train: y_pitch %d < c %d", y_pitch, c);
    if ((moving_mean == nullptr) != (moving_var == nullptr)) return fail(T2I_ERR_BAD_ARG, "bn_apply_train: moving pair");
    int CG;
    dim3 grid;
    if (np != 1 && np != 2) return fail(T2I_ERR_BAD_ARG, "bn_apply_train: np must be 1 or 2");
    const bool has_res = residual != nullptr;
    rowwise_geometry(rows, c, 256, &CG, &grid, (np == 1 && !has_res) ? 8 : 4);
    const long long n = stat_rows > 0 ? stat_rows : rows;     // values per channel behind the sums
    const float bessel = n > 1 ? (float)n / (float)(n - 1) : 1.f;
#define T2I_BN_APPLY(NP, RES)                                                                                            \
    launch_ew(bn_apply_train_kernel<NP, RES>, dim3(grid), dim3(256), 0, STREAM, static_cast<const bf16*>(x), x_ps, sums, \
              1.f / (float)n, eps, gamma, beta, static_cast<const bf16*>(residual), r_ps, static_cast<bf16*>(y), y_ps,  \
              rows, c, relu, CG, mean, rstd, var, moving_mean, moving_var, decay, bessel, y_pitch, affine_scale)
    if (np == 1) {
        if (has_res) T2I_BN_APPLY(1, true); else T2I_BN_APPLY(1, false);
    } else {
        if (has_res) T2I_BN_APPLY(2, true); else T2I_BN_APPLY(2, false);
    }
#undef T2I_BN_APPLY
    return check_launch("bn_apply_train");
}
extern "C" int t2i_bn_bwd_fused(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* mean,
                                const float* rstd, const float* gamma, const float* dot, const float* dbeta,
                                float* dgamma, float* dbeta_out, float out_scale, int dot_normalised, void* dx,
                                long long dx_ps, float* dx_sum, int np, long long rows, int c, long long stat_rows,
                                int dy_pitch, float affine_scale, void* stream) {
    if (c % 8 || dy_pitch % 8) return fail(T2I_ERR_BAD_ARG, "bn_bwd_fused: c and dy_pitch must be multiples of 8");
    if (dy_pitch == 0) dy_pitch = c;
    int CG;
    dim3 grid;
    if (np != 1 && np != 2) return fail(T2I_ERR_BAD_ARG, "bn_bwd_fused: np must be 1 or 2");
    rowwise_geometry(rows, c, 256, &CG, &grid, 8);
    const long long n = stat_rows > 0 ? stat_rows : rows;
#define T2I_BN_BWD(NP)                                                                                                   \
    launch_ew(bn_bwd_fused_kernel<NP>, dim3(grid), dim3(256), 0, STREAM, static_cast<const bf16*>(dy), dy_ps,            \
              static_cast<const bf16*>(x), x_ps, mean, rstd, gamma, dot, dbeta, dgamma, dbeta_out, out_scale,           \
              dot_normalised, static_cast<bf16*>(dx), dx_ps, dx_sum, rows, c, 1.f / (float)n, CG, dy_pitch, affine_scale)
    if (np == 1) T2I_BN_BWD(1); else T2I_BN_BWD(2);
#undef T2I_BN_BWD
    return check_launch("bn_bwd_fused");
}
extern "C" int t2i_bn_update_moving(float* mm, float* mv, const float* mean, const float* var, long long rows, int c,
                                    float decay, void* stream) {
    const float bessel = rows > 1 ? (float)rows / (float)(rows - 1) : 1.f;
    launch_ew(bn_update_moving_kernel, dim3(ceil_div(c, 256)), dim3(256), 0, STREAM, mm, mv, mean, var, c, decay, bessel);
    return check_launch("bn_update_moving");
}
extern "C" int t2i_act_bwd(const void* dy, long long dy_ps, const void* y, long long y_ps, void* dst, long long dst_ps,
                           int np, long long n, int mask_kind, void* stream) {
    if (n % 8) return fail(T2I_ERR_BAD_ARG, "act_bwd: n must be a multiple of 8");
    const float neg = (mask_kind == T2I_MASK_LRELU) ? 0.2f : 0.f;
    launch_ew(act_bwd_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, STREAM, static_cast<const bf16*>(dy), dy_ps, static_cast<const bf16*>(y),
                                                            y_ps, static_cast<bf16*>(dst), dst_ps, np, n / 8, neg);
    return check_launch("act_bwd");
}
extern "C" int t2i_embed_tile(const void* e, long long e_ps, void* cat, long long cat_ps, int np, int s, int c, int pitch,
                              int coff, int hw, void* stream) {
    if (c % 8 || pitch % 8 || coff % 8) return fail(T2I_ERR_BAD_ARG, "embed_tile: channels must be multiples of 8");
    launch_ew(embed_tile_kernel, dim3(grid_for((long long)s * hw * (c / 8), 256)), dim3(256), 0, STREAM, 
        static_cast<const bf16*>(e), e_ps, static_cast<bf16*>(cat), cat_ps, np, s, c, pitch, coff, hw);
    return check_launch("embed_tile");
}
extern "C" int t2i_embed_reduce(const void* dcat, long long dcat_ps, void* de, long long de_ps, int np, int s, int c,
                                int pitch, int coff, int hw, void* stream) {
    if (c % 8 || pitch % 8 || coff % 8) return fail(T2I_ERR_BAD_ARG, "embed_reduce: channels must be multiples of 8");
    launch_ew(embed_reduce_kernel, dim3(grid_for((long long)s * (c / 8), 128)), dim3(128), 0, STREAM, 
        static_cast<const bf16*>(dcat), dcat_ps, static_cast<bf16*>(de), de_ps, np, s, c, pitch, coff, hw);
    return check_launch("embed_reduce");
}
extern "C" int t2i_dout_fwd(const void* a, long long a_ps, int np, const float* w, const float* b, float* logit, int s,
                            int k, void* stream) {
    if (k % 8) return fail(T2I_ERR_BAD_ARG, "dout_fwd: k must be a multiple of 8");
    launch_ew(dout_fwd_kernel, dim3(s), dim3(256), 0, STREAM, static_cast<const bf16*>(a), a_ps, np, w, b, logit, k);
    return check_launch("dout_fwd");
}
extern "C" int t2i_dout_bwd_data(const void* a, long long a_ps, int np, const float* w, const float* seed, void* da,
                                 long long da_ps, int s, int k, void* stream) {
    if (k % 8) return fail(T2I_ERR_BAD_ARG, "dout_bwd_data: k must be a multiple of 8");
    launch_ew(dout_bwd_data_kernel, dim3(grid_for((long long)s * (k / 8), 256)), dim3(256), 0, STREAM, 
        static_cast<const bf16*>(a), a_ps, np, w, seed, static_cast<bf16*>(da), da_ps, s, k);
    return check_launch("dout_bwd_data");
}
extern "C" int t2i_dout_bwd_weight(const void* a, long long a_ps, int np, const float* seed, float* dw, float* db, int s,
                                   int s_bias, int k, void* stream) {
    if (k % 8) return fail(T2I_ERR_BAD_ARG, "dout_bwd_weight: k must be a multiple of 8");
    const int gx = ceil_div(k / 8, 128);
    int gy = s < 32 ? s : 32;
    launch_ew(dout_bwd_weight_kernel, dim3(dim3(gx, gy)), dim3(128), 0, STREAM, static_cast<const bf16*>(a), a_ps, np, seed, dw, s, k);
    int rc = check_launch("dout_bwd_weight");
    if (rc != T2I_OK) return rc;
    if (db != nullptr && s_bias > 0) {
        launch_ew(seed_sum_kernel, dim3(1), dim3(256), 0, STREAM, seed, s_bias, db);
        rc = check_launch("seed_sum");
    }
    return rc;
}
extern "C" int t2i_gp_interp(const float* g, const float* x, const float* eps, float* xhat, int n, int per_sample,
                             void* stream) {
    if (per_sample % 4) return fail(T2I_ERR_BAD_ARG, "gp_interp: per_sample must be a multiple of 4");
    const long long n4 = (long long)n * per_sample / 4;
    launch_ew(gp_interp_kernel, dim3(grid_for(n4, 256)), dim3(256), 0, STREAM, g, x, eps, xhat, n4, per_sample / 4);
    return check_launch("gp_interp");
}
extern "C" int t2i_gp_penalty(const float* grad, int n, int per_sample, float weight, float inv_global_batch,
                              float* slope, float* coef, float* pen_sum, void* stream) {
    if (per_sample % 4) return fail(T2I_ERR_BAD_ARG, "gp_penalty: per_sample must be a multiple of 4");
    launch_ew(gp_penalty_kernel, dim3(n), dim3(256), 0, STREAM, grad, per_sample, weight, inv_global_batch, slope, coef, pen_sum);
    return check_launch("gp_penalty");
}
extern "C" int t2i_ca_fwd(const float* ms, const float* z, const float* tn_eps, void* zc, long long zc_ps,
                          int np, int b, int z_dim, int ce, float* kl_sum, void* stream) {
    launch_ew(ca_fwd_kernel, dim3(grid_for((long long)b * (z_dim + ce), 256, 1)), dim3(256), 0, STREAM,
        ms, z, tn_eps, static_cast<bf16*>(zc), zc_ps, np, b, z_dim, ce, kl_sum);
    return check_launch("ca_fwd");
}
extern "C" int t2i_ca_bwd(const float* ms, const void* dzc, long long dzc_ps, const float* tn_eps,
                          void* dms, long long dms_ps, int np, int b, int z_dim, int ce, float kl_scale, void* stream) {
    launch_ew(ca_bwd_kernel, dim3(grid_for((long long)b * ce, 256, 1)), dim3(256), 0, STREAM,
        ms, static_cast<const bf16*>(dzc), dzc_ps, tn_eps, static_cast<bf16*>(dms), dms_ps,
        np, b, z_dim, ce, kl_scale);
    return check_launch("ca_bwd");
}
extern "C" int t2i_d_seeds(const float* kt, float* seed, int b, float inv_global_batch, void* stream) {
    launch_ew(d_seeds_kernel, dim3(ceil_div(4 * b, 256)), dim3(256), 0, STREAM, kt, seed, b, inv_global_batch);
    return check_launch("d_seeds");
}
extern "C" int t2i_d_sums(const float* logit, int b, float* sums, void* stream) {
    launch_ew(d_sums_kernel, dim3(1), dim3(256), 0, STREAM, logit, b, sums);
    return check_launch("d_sums");
}
extern "C" int t2i_d_scalars(const float* sums, float* kt, float* scalars, int global_batch, float gp_weight, float kt_lr,
                             void* stream) {
    launch_ew(d_scalars_kernel, dim3(1), dim3(32), 0, STREAM, sums, kt, scalars, 1.f / (float)global_batch, gp_weight, kt_lr);
    return check_launch("d_scalars");
}
extern "C" int t2i_ce_seeds(const float* logit, int n, float label, float weight, float inv_global_batch, float* seed,
                            float* loss_sum, void* stream) {
    launch_ew(ce_seeds_kernel, dim3(1), dim3(256), 0, STREAM, logit, n, label, weight, inv_global_batch, seed, loss_sum);
    return check_launch("ce_seeds");
}
extern "C" int t2i_s1_scalars(const float* sums, float* scalars, int global_batch, int ce, float alpha, float kl_coeff,
                              int which, void* stream) {
    launch_ew(s1_scalars_kernel, dim3(1), dim3(32), 0, STREAM, sums, scalars, 1.f / (float)global_batch,
              1.f / ((float)global_batch * (float)ce), alpha, kl_coeff, which);
    return check_launch("s1_scalars");
}
extern "C" int t2i_g_sums(const float* logit_fake, int b, float* sums, void* stream) {
    launch_ew(g_sums_kernel, dim3(1), dim3(256), 0, STREAM, logit_fake, b, sums);
    return check_launch("g_sums");
}
extern "C" int t2i_g_scalars(const float* sums, float* scalars, int global_batch, int ce, float kl_coeff, void* stream) {
    launch_ew(g_scalars_kernel, dim3(1), dim3(32), 0, STREAM, sums, scalars, 1.f / (float)global_batch,
                                           1.f / ((float)global_batch * (float)ce), kl_coeff);
    return check_launch("g_scalars");
}
extern "C" int t2i_pack_weight(const float* w, int taps, int cout, int cin, void* fwd, long long fwd_ps, void* bwd,
                               long long bwd_ps, int np, void* stream) {
    dim3 grid(ceil_div(cin, 32), ceil_div(cout, 32), taps);
    launch_ew(pack_weight_kernel, grid, dim3(32, 8), 0, STREAM, w, cout, cin, static_cast<bf16*>(fwd), fwd_ps,
                                                        static_cast<bf16*>(bwd), bwd_ps, np);
    return check_launch("pack_weight");
}
extern "C" int t2i_adam_tf(float* theta, const float* grad, float* m, float* v, long long n, const float* lr_t_dev,
                           float beta1, float beta2, float eps, float grad_scale, void* packed, long long packed_ps, int np,
                           void* stream) {
    launch_ew(adam_tf_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, STREAM, theta, grad, m, v, n, lr_t_dev, beta1, beta2, eps,
                                                               grad_scale, static_cast<bf16*>(packed), packed_ps, np);
    return check_launch("adam_tf");
}
