// ln2d.cu — LayerNorm over the channel dimension of a channel-first (B, C, H, W) tensor, forward and backward.
//
// Replaces LayerNorm2d (basicsr/vmamba/models/vmamba.py:58-63): permute to (B, H, W, C), F.layer_norm, permute back. On a
// channel-first network every VSSBlock calls it three times (norm, SS2D.out_norm, norm2: vmamba.py:1239-1262), and the eager
// form costs two layout copies around a kernel written for long rows: for C = 40..160 and 32768 pixels (the 8 x 128 x 128
// training batch of BASELINE configs[4]) PyTorch's gamma/beta gradient kernel alone takes 90 us per call, 15 % of the
// training step's GPU time (profiles/r02_train_step.md).
//
// Here the tensor stays channel-first. One thread owns one pixel: its C values are C coalesced rows apart, so every load and
// store of a warp is a contiguous 128-byte line, and the three passes over the pixel's channels (mean, variance, output)
// re-read lines the CTA just brought into L1. The parameter gradients are a separate streaming reduction organised by
// channel row. All three kernels are bound by launch latency and DRAM bytes: 2 / 4 / 2 tensor passes.
#include "bem_kernels.h"

namespace bem {

constexpr int kLnThreads = 128;

__global__ void __launch_bounds__(kLnThreads) ln2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, float* __restrict__ y,
                                                              float* __restrict__ mean, float* __restrict__ rstd, int64_t npix, int C,
                                                              int64_t HW, float eps) {
    pdl_trigger();
    pdl_wait();
    const int64_t p = (int64_t)blockIdx.x * kLnThreads + threadIdx.x;
    if (p >= npix) return;
    const int64_t bb = p / HW, q = p - bb * HW;
    const float* xp = x + bb * C * HW + q;
    float* yp = y + bb * C * HW + q;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
        s0 += xp[(int64_t)c * HW];
        s1 += xp[(int64_t)(c + 1) * HW];
        s2 += xp[(int64_t)(c + 2) * HW];
        s3 += xp[(int64_t)(c + 3) * HW];
    }
    for (; c < C; ++c) s0 += xp[(int64_t)c * HW];
    const float inv = 1.f / (float)C;
    const float m = ((s0 + s1) + (s2 + s3)) * inv;
    s0 = s1 = s2 = s3 = 0.f;
    for (c = 0; c + 4 <= C; c += 4) {
        const float d0 = xp[(int64_t)c * HW] - m, d1 = xp[(int64_t)(c + 1) * HW] - m;
        const float d2 = xp[(int64_t)(c + 2) * HW] - m, d3 = xp[(int64_t)(c + 3) * HW] - m;
        s0 = fmaf(d0, d0, s0);
        s1 = fmaf(d1, d1, s1);
        s2 = fmaf(d2, d2, s2);
        s3 = fmaf(d3, d3, s3);
    }
    for (; c < C; ++c) {
        const float d0 = xp[(int64_t)c * HW] - m;
        s0 = fmaf(d0, d0, s0);
    }
    const float r = rsqrtf(((s0 + s1) + (s2 + s3)) * inv + eps);
    if (mean) {
        mean[p] = m;
        rstd[p] = r;
    }
#pragma unroll 4
    for (c = 0; c < C; ++c) {
        const float xh = (xp[(int64_t)c * HW] - m) * r;
        yp[(int64_t)c * HW] = w ? fmaf(xh, w[c], b ? b[c] : 0.f) : xh;
    }
}

// dx = rstd * (g w - mean_c(g w) - xhat * mean_c(g w xhat))
__global__ void __launch_bounds__(kLnThreads) ln2d_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 const float* __restrict__ w, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, float* __restrict__ dx, int64_t npix,
                                                                 int C, int64_t HW) {
    pdl_trigger();
    pdl_wait();
    const int64_t p = (int64_t)blockIdx.x * kLnThreads + threadIdx.x;
    if (p >= npix) return;
    const int64_t bb = p / HW, q = p - bb * HW;
    const float* xp = x + bb * C * HW + q;
    const float* gp = dy + bb * C * HW + q;
    float* dp = dx + bb * C * HW + q;
    const float m = mean[p], r = rstd[p];
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    int c = 0;
    for (; c + 2 <= C; c += 2) {
        const float g0 = gp[(int64_t)c * HW] * (w ? w[c] : 1.f), g1 = gp[(int64_t)(c + 1) * HW] * (w ? w[c + 1] : 1.f);
        const float h0 = (xp[(int64_t)c * HW] - m) * r, h1 = (xp[(int64_t)(c + 1) * HW] - m) * r;
        a0 += g0;
        a1 += g1;
        b0 = fmaf(g0, h0, b0);
        b1 = fmaf(g1, h1, b1);
    }
    for (; c < C; ++c) {
        const float g0 = gp[(int64_t)c * HW] * (w ? w[c] : 1.f);
        a0 += g0;
        b0 = fmaf(g0, (xp[(int64_t)c * HW] - m) * r, b0);
    }
    const float inv = 1.f / (float)C;
    const float sa = (a0 + a1) * inv, sb = (b0 + b1) * inv;
#pragma unroll 4
    for (c = 0; c < C; ++c) {
        const float g = gp[(int64_t)c * HW] * (w ? w[c] : 1.f);
        const float xh = (xp[(int64_t)c * HW] - m) * r;
        dp[(int64_t)c * HW] = r * (g - sa - xh * sb);
    }
}

// dweight[c] = sum over pixels of dy * xhat, dbias[c] = sum of dy. One CTA per (channel, batch image, split of the row).
__global__ void __launch_bounds__(256) ln2d_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ dw, float* __restrict__ db, int C, int64_t HW,
                                                             int splits) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x, bb = blockIdx.y, sp = blockIdx.z;
    const int64_t per = (HW + splits - 1) / splits;
    const int64_t q0 = sp * per, q1 = min(HW, q0 + per);
    const float* gp = dy + ((int64_t)bb * C + c) * HW;
    const float* xp = x + ((int64_t)bb * C + c) * HW;
    const float* mp = mean + (int64_t)bb * HW;
    const float* rp = rstd + (int64_t)bb * HW;
    float s1 = 0.f, s2 = 0.f;
    for (int64_t q = q0 + threadIdx.x; q < q1; q += 256) {
        const float g = gp[q];
        s1 = fmaf(g, (xp[q] - mp[q]) * rp[q], s1);
        s2 += g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ float sh[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[0][warp] = s1;
        sh[1][warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            t1 += sh[0][i];
            t2 += sh[1][i];
        }
        if (dw) atomicAdd(dw + c, t1);
        if (db) atomicAdd(db + c, t2);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Tiled forms for C <= 160 (every BEM / VMamba width): 8 threads per pixel, each holding KC = ceil(C / 8) channels in
// registers — one read of the tensor per pass, eight times the loads in flight of the one-thread-per-pixel kernels above, and
// the parameter gradients fall out of the same pass (warp = 32 pixels of one channel slice: shuffle-reduce, one atomic per
// channel per 128 pixels). Block = 32 pixels x 8 slices.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLnSlices = 8;

template <int KC>
__global__ void __launch_bounds__(256) ln2d_fwd_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, float* __restrict__ y, float* __restrict__ mean,
                                                             float* __restrict__ rstd, int64_t npix, int C, int64_t HW, float eps) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[2][kLnSlices][32];
    const int px = threadIdx.x & 31, cs = threadIdx.x >> 5;
    const int64_t p = (int64_t)blockIdx.x * 32 + px;
    const bool live = p < npix;
    const int64_t bb = live ? p / HW : 0, q = live ? p - bb * HW : 0;
    const float* xp = x + bb * C * HW + q;
    float v[KC];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = cs + kLnSlices * k;
        v[k] = (live && c < C) ? xp[(int64_t)c * HW] : 0.f;
        s += v[k];
    }
    red[0][cs][px] = s;
    __syncthreads();
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < kLnSlices; ++i) m += red[0][i][px];
    m *= 1.f / (float)C;
    float d2 = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const float d = (cs + kLnSlices * k < C) ? v[k] - m : 0.f;
        v[k] = d;
        d2 = fmaf(d, d, d2);
    }
    red[1][cs][px] = d2;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kLnSlices; ++i) var += red[1][i][px];
    const float r = rsqrtf(var * (1.f / (float)C) + eps);
    if (!live) return;
    if (mean && cs == 0) {
        mean[p] = m;
        rstd[p] = r;
    }
    float* yp = y + bb * C * HW + q;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = cs + kLnSlices * k;
        if (c < C) {
            const float xh = v[k] * r;
            yp[(int64_t)c * HW] = w ? fmaf(xh, w[c], b ? b[c] : 0.f) : xh;
        }
    }
}

template <int KC>
__global__ void __launch_bounds__(256) ln2d_bwd_tiled_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ w, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ dx, float* __restrict__ dw,
                                                             float* __restrict__ db, int64_t npix, int C, int64_t HW, int groups) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[2][kLnSlices][32];
    const int px = threadIdx.x & 31, cs = threadIdx.x >> 5;
    float wk[KC], aw[KC], ab[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = cs + kLnSlices * k;
        wk[k] = (w && c < C) ? w[c] : 1.f;
        aw[k] = ab[k] = 0.f;
    }
    const float inv = 1.f / (float)C;
    for (int gi = 0; gi < groups; ++gi) {
        const int64_t p = ((int64_t)blockIdx.x * groups + gi) * 32 + px;
        const bool live = p < npix;
        const int64_t bb = live ? p / HW : 0, q = live ? p - bb * HW : 0;
        const float* xp = x + bb * C * HW + q;
        const float* gp = dy + bb * C * HW + q;
        const float m = live ? mean[p] : 0.f, r = live ? rstd[p] : 0.f;
        float g[KC], xh[KC];
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = cs + kLnSlices * k;
            const bool on = live && c < C;
            g[k] = on ? gp[(int64_t)c * HW] : 0.f;
            xh[k] = on ? (xp[(int64_t)c * HW] - m) * r : 0.f;
            aw[k] = fmaf(g[k], xh[k], aw[k]);
            ab[k] += g[k];
            const float gw = g[k] * wk[k];
            sa += gw;
            sb = fmaf(gw, xh[k], sb);
        }
        if (dx) {
            red[0][cs][px] = sa;
            red[1][cs][px] = sb;
            __syncthreads();
            float ta = 0.f, tb = 0.f;
#pragma unroll
            for (int i = 0; i < kLnSlices; ++i) {
                ta += red[0][i][px];
                tb += red[1][i][px];
            }
            ta *= inv;
            tb *= inv;
            if (live) {
                float* dp = dx + bb * C * HW + q;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int c = cs + kLnSlices * k;
                    if (c < C) dp[(int64_t)c * HW] = r * (g[k] * wk[k] - ta - xh[k] * tb);
                }
            }
            __syncthreads();
        }
    }
    if (dw || db) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            float a = aw[k], bsum = ab[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
            }
            const int c = cs + kLnSlices * k;
            if (px == 0 && c < C) {
                if (dw) atomicAdd(dw + c, a);
                if (db) atomicAdd(db + c, bsum);
            }
        }
    }
}

}  // namespace bem

using namespace bem;

extern "C" int bem_layernorm2d_fwd(const float* x, const float* weight, const float* bias, float* y, float* mean, float* rstd,
                                   int32_t batch, int32_t channels, int64_t hw, float eps, void* stream_) {
    if (!x || !y || batch <= 0 || channels <= 0 || hw <= 0 || (mean == nullptr) != (rstd == nullptr) || (bias && !weight)) return BEM_ERR_BAD_ARG;
    const int64_t npix = (int64_t)batch * hw;
    const int64_t blocks = (npix + kLnThreads - 1) / kLnThreads;
    if (blocks > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t tblocks = (npix + 31) / 32;
    if (channels <= 160 && tblocks <= 0x7fffffff) {
        const dim3 grid((unsigned)tblocks), block(256);
        if (channels <= 40) launch_pdl(ln2d_fwd_tiled_kernel<5>, grid, block, 0, stream, x, weight, bias, y, mean, rstd, npix, (int)channels, hw, eps);
        else if (channels <= 80) launch_pdl(ln2d_fwd_tiled_kernel<10>, grid, block, 0, stream, x, weight, bias, y, mean, rstd, npix, (int)channels, hw, eps);
        else launch_pdl(ln2d_fwd_tiled_kernel<20>, grid, block, 0, stream, x, weight, bias, y, mean, rstd, npix, (int)channels, hw, eps);
        return (int)cudaGetLastError();
    }
    launch_pdl(ln2d_fwd_kernel, dim3((unsigned)blocks), dim3(kLnThreads), 0, stream, x, weight, bias, y, mean, rstd, npix,
               (int)channels, hw, eps);
    return (int)cudaGetLastError();
}

extern "C" int bem_layernorm2d_bwd(const float* dy, const float* x, const float* weight, const float* mean, const float* rstd, float* dx,
                                   float* dweight, float* dbias, int32_t batch, int32_t channels, int64_t hw, void* stream_) {
    if (!dy || !x || !mean || !rstd || batch <= 0 || channels <= 0 || hw <= 0) return BEM_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t npix = (int64_t)batch * hw;
    const int64_t blocks = (npix + kLnThreads - 1) / kLnThreads;
    if (blocks > 0x7fffffff || batch > 65535) return BEM_ERR_UNSUPPORTED;
    if (channels <= 160) {
        // Pixel groups (of 32) per block: every block ends with one atomic per channel onto the same C addresses, and those
        // serialise in L2 (1024 blocks x 40 channels took 50 us at 32768 pixels) - so no more blocks than fill the machine twice.
        int groups = (int)((npix / 32 + 2LL * device_sm_count() - 1) / (2LL * device_sm_count()));
        groups = groups < 1 ? 1 : (groups > 64 ? 64 : groups);
        const int64_t tblocks = (npix + 32 * groups - 1) / (32 * groups);
        const dim3 grid((unsigned)tblocks), block(256);
        if (channels <= 40) launch_pdl(ln2d_bwd_tiled_kernel<5>, grid, block, 0, stream, dy, x, weight, mean, rstd, dx, dweight, dbias, npix, (int)channels, hw, groups);
        else if (channels <= 80) launch_pdl(ln2d_bwd_tiled_kernel<10>, grid, block, 0, stream, dy, x, weight, mean, rstd, dx, dweight, dbias, npix, (int)channels, hw, groups);
        else launch_pdl(ln2d_bwd_tiled_kernel<20>, grid, block, 0, stream, dy, x, weight, mean, rstd, dx, dweight, dbias, npix, (int)channels, hw, groups);
        return (int)cudaGetLastError();
    }
    if (dx) launch_pdl(ln2d_bwd_dx_kernel, dim3((unsigned)blocks), dim3(kLnThreads), 0, stream, dy, x, weight, mean, rstd, dx, npix, (int)channels, hw);
    if (dweight || dbias) {
        // enough CTAs to fill the machine: split long rows (dweight / dbias are accumulated with atomics, zero-filled by the caller)
        int splits = 1;
        while ((int64_t)channels * batch * splits < 2LL * device_sm_count() && hw / (splits * 2) >= 2048 && splits < 64) splits *= 2;
        launch_pdl(ln2d_bwd_param_kernel, dim3((unsigned)channels, (unsigned)batch, (unsigned)splits), dim3(256), 0, stream, dy, x, mean, rstd,
                   dweight, dbias, (int)channels, hw, splits);
    }
    return (int)cudaGetLastError();
}
