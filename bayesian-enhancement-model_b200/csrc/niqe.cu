// niqe.cu — the per-pixel part of the device-resident NIQE scorer (bem_b200/niqe.py), fp64 like the reference's numpy code.
//
// Replaces, for a batch of images at once, the inner loops of basicsr/metrics/niqe.py:
//   niqe_mscn_kernel        :104-108  mu = convolve(img, w7x7, 'nearest'); sigma = sqrt(|convolve(img^2, w) - mu^2|);
//                                      img_normalized = (img - mu) / (sigma + 1)
//   niqe_block_stats_kernel :110-116, 41-60, 13-38  per 96 x 96 (48 x 48 at the second scale) block the moments the AGGD fits
//                                      need, for the block itself and for its products with four circular shifts
//                                      (np.roll(block, s, axis=(0, 1)), s in (0,1), (1,0), (1,1), (1,-1)):
//                                      sum_{v<0} v^2, #{v<0}, sum_{v>0} v^2, #{v>0}, sum |v|, sum v^2
// Everything after that (moment matching against the gamma table, the 36-d Gaussian fit, the pseudo-inverse) is a few
// kilobytes per image and stays in bem_b200/niqe.py.
#include <cuda_runtime.h>

#include "bem_kernels.h"

namespace bem {

__global__ void __launch_bounds__(256) niqe_mscn_kernel(const float* __restrict__ img, const double* __restrict__ win, double* __restrict__ out,
                                                         int S, int H, int W) {
    pdl_trigger();
    pdl_wait();
    __shared__ double w[49];
    if (threadIdx.x < 49) w[threadIdx.x] = win[threadIdx.x];
    __syncthreads();
    const int64_t total = (int64_t)S * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int64_t t = i / W;
        const int y = (int)(t % H);
        const float* plane = img + (t / H) * (int64_t)H * W;
        double mu = 0.0, m2 = 0.0;
        // scipy.ndimage.convolve flips the window; the NIQE window is symmetric, so correlation order is kept. 'nearest' =
        // clamped coordinates.
#pragma unroll
        for (int dy = -3; dy <= 3; ++dy) {
            const float* row = plane + (int64_t)min(max(y + dy, 0), H - 1) * W;
#pragma unroll
            for (int dx = -3; dx <= 3; ++dx) {
                const double v = (double)row[min(max(x + dx, 0), W - 1)];
                const double ww = w[(3 - dy) * 7 + (3 - dx)];
                mu = fma(ww, v, mu);
                m2 = fma(ww, v * v, m2);
            }
        }
        const double c = (double)plane[(int64_t)y * W + x];
        out[i] = (c - mu) / (sqrt(fabs(m2 - mu * mu)) + 1.0);
    }
}

// grid = S * nbh * nbw blocks of 256 threads; out: (S, nbh * nbw, 5, 6) doubles
__global__ void __launch_bounds__(256) niqe_block_stats_kernel(const double* __restrict__ nrm, double* __restrict__ out, int S, int H, int W,
                                                               int bs) {
    pdl_trigger();
    pdl_wait();
    const int nbw = W / bs, nbh = H / bs;
    const int blk = blockIdx.x % (nbh * nbw), s = blockIdx.x / (nbh * nbw);
    const int bh = blk / nbw, bw = blk - bh * nbw;
    const double* base = nrm + (int64_t)s * H * W + (int64_t)bh * bs * W + (int64_t)bw * bs;
    double acc[5][6];
#pragma unroll
    for (int m = 0; m < 5; ++m)
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[m][k] = 0.0;
    for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) {
        const int i = e / bs, j = e - i * bs;
        const int im = i == 0 ? bs - 1 : i - 1, jm = j == 0 ? bs - 1 : j - 1, jp = j == bs - 1 ? 0 : j + 1;
        const double x = base[(int64_t)i * W + j];
        double v[5];
        v[0] = x;
        v[1] = x * base[(int64_t)i * W + jm];      // roll (0, 1):  rolled[i][j] = block[i][j - 1]
        v[2] = x * base[(int64_t)im * W + j];      // roll (1, 0)
        v[3] = x * base[(int64_t)im * W + jm];     // roll (1, 1)
        v[4] = x * base[(int64_t)im * W + jp];     // roll (1, -1): rolled[i][j] = block[i - 1][j + 1]
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            const double q = v[m], q2 = q * q;
            if (q < 0.0) {
                acc[m][0] += q2;
                acc[m][1] += 1.0;
            } else if (q > 0.0) {
                acc[m][2] += q2;
                acc[m][3] += 1.0;
            }
            acc[m][4] += fabs(q);
            acc[m][5] += q2;
        }
    }
    __shared__ double red[8][30];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int m = 0; m < 5; ++m)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double a = acc[m][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[warp][m * 6 + k] = a;
        }
    __syncthreads();
    if (threadIdx.x < 30) {
        double a = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) a += red[w8][threadIdx.x];     // fixed order: reproducible
        out[(int64_t)blockIdx.x * 30 + threadIdx.x] = a;
    }
}

}  // namespace bem

extern "C" {

int bem_niqe_mscn(const float* img, const double* window7x7, double* out, int32_t n_images, int32_t H, int32_t W, void* stream) {
    if (!img || !window7x7 || !out || n_images <= 0 || H <= 0 || W <= 0) return BEM_ERR_BAD_ARG;
    const int64_t total = (int64_t)n_images * H * W;
    const int grid = (int)((total + 255) / 256 < (int64_t)bem::device_sm_count() * 16 ? (total + 255) / 256 : (int64_t)bem::device_sm_count() * 16);
    bem::launch_pdl(bem::niqe_mscn_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, img, window7x7, out, (int)n_images, (int)H, (int)W);
    return (int)cudaGetLastError();
}

int bem_niqe_block_stats(const double* normalized, double* out, int32_t n_images, int32_t H, int32_t W, int32_t block, void* stream) {
    if (!normalized || !out || n_images <= 0 || block <= 1 || H < block || W < block || H % block || W % block) return BEM_ERR_BAD_ARG;
    const int64_t blocks = (int64_t)n_images * (H / block) * (W / block);
    if (blocks > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    bem::launch_pdl(bem::niqe_block_stats_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, normalized, out, (int)n_images,
                    (int)H, (int)W, (int)block);
    return (int)cudaGetLastError();
}

}  // extern "C"
