// bayes.cu — reparameterised Bayesian layers for sm_100a.
//
// Replaces the eager chain of Conv2dReparameterization / Linear2dReparameterization._forward_uncertain
// (basicsr/bayesian/conv.py:106-114, linear.py:82-90): sigma = log1p(exp(rho)), eps.normal_(), w = mu + sigma * eps
// (4-6 tiny elementwise launches per layer per sample) followed by a library convolution.
//   bem_bayes_sample     one launch for all S samples of a tensor; eps either given (parity with the reference) or
//                        generated in-kernel by a counter-based Philox4x32-10 + Box-Muller stream keyed
//                        (seed, tensor id, sample) so that results do not depend on how samples are sharded over GPUs
//   bem_bayes_pointwise  S-batched 1x1 convolution, per-sample weights; with mu/rho/eps given the sample step is fused
//                        into the weight-tile load (the sampled weights never exist in HBM)
//   bem_bayes_depthwise  S-batched depthwise 3x3, per-sample weights
// This file holds the fp32 CUDA-core kernels (bit-faithful fp32 accumulate: the 1e-5 parity tier);
// the tcgen05 tensor-core kernel of the pointwise contraction (the default) lives in bayes_tc.cu.
#include <cstdlib>

#include "bem_kernels.h"

namespace bem {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), restated in oracle/philox.py
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// 4 standard normals from one Philox block: Box-Muller on (x0,x1) and (x2,x3); u = (x >> 8 + 0.5) * 2^-24 in (0,1)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t stream_id, uint64_t sample, uint64_t block, float (&n)[4]) {
    const uint4 ctr = make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)sample, (uint32_t)stream_id);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 r = philox4x32_10(ctr, key);
    const float k = 5.9604644775390625e-08f;   // 2^-24
    const float u0 = ((float)(r.x >> 8) + 0.5f) * k, u1 = ((float)(r.y >> 8) + 0.5f) * k;
    const float u2 = ((float)(r.z >> 8) + 0.5f) * k, u3 = ((float)(r.w >> 8) + 0.5f) * k;
    const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
    const float ta = 6.283185307179586f * u1, tb = 6.283185307179586f * u3;
    n[0] = ra * cosf(ta);
    n[1] = ra * sinf(ta);
    n[2] = rb * cosf(tb);
    n[3] = rb * sinf(tb);
}

__device__ __forceinline__ float sigma_of_rho(float rho) { return log1pf(expf(rho)); }   // conv.py:106

__global__ void __launch_bounds__(256) bayes_sample_kernel(const BemBayesSampleParams p) {
    pdl_trigger();
    pdl_wait();
    const int64_t nblk = (p.numel + 3) / 4;
    const int64_t total = nblk * p.n_samples;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / nblk, blk = t - s * nblk;
        float e[4];
        if (p.rho && !p.eps) philox_normal4(p.seed, p.stream_id, (uint64_t)(p.sample0 + s), (uint64_t)blk, e);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = blk * 4 + k;
            if (i >= p.numel) break;
            float w = p.mu[i];
            if (p.rho) {
                const float ev = p.eps ? p.eps[s * p.numel + i] : e[k];
                w = fmaf(sigma_of_rho(p.rho[i]), ev, w);
                if (p.eps_out) p.eps_out[s * p.numel + i] = ev;
            }
            p.w[s * p.numel + i] = w;
        }
    }
}

// every Bayesian tensor of a network in one launch (one Monte-Carlo draw); same numbers as bayes_sample_kernel
__global__ void __launch_bounds__(256) bayes_sample_batched_kernel(const BemBayesSampleBatchedParams p) {
    pdl_trigger();
    pdl_wait();
    const int e = p.blocks[2 * blockIdx.x];
    const int64_t blk = (int64_t)p.blocks[2 * blockIdx.x + 1] + threadIdx.x;
    const BemBayesSampleEntry en = p.entries[e];
    if (blk * 4 >= en.numel) return;
    const int64_t sample = p.sample0 + (p.sample0_dev ? *p.sample0_dev : 0);
    float n[4];
    philox_normal4(p.seed, (uint64_t)en.stream_id, (uint64_t)sample, (uint64_t)blk, n);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = blk * 4 + k;
        if (i < en.numel) en.w[i] = fmaf(sigma_of_rho(en.rho[i]), n[k], en.mu[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// S-batched pointwise (1x1) convolution, fp32 CUDA-core tiles.
//   out[img, co, p] = sum_ci W[s(img)][co][ci] * x[img][ci][p] + bias[s(img)][co]
// CTA tile: 64 output channels x 256 pixels, K step 16. 256 threads as (32 pixel lanes) x (8 channel groups):
// each thread owns 8 channels x 8 pixels. Weight reads are warp-broadcasts, x reads conflict-free LDS.128.
// ------------------------------------------------------------------------------------------------
constexpr int PW_BM = 64, PW_BN = 256, PW_BK = 16;

__global__ void __launch_bounds__(256) bayes_pointwise_kernel(const BemBayesPointwiseParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float sW[PW_BK][PW_BM + 4];   // +4: the transposing store is 2-way instead of 16-way conflicted
    __shared__ __align__(16) float sX[PW_BK][PW_BN];
    const int img = blockIdx.z;
    const int imgs_per_sample = p.batch / p.n_samples;
    const int s = p.n_samples > 1 ? (p.sample_interleave ? img % p.n_samples : img / imgs_per_sample) : 0;
    const int co0 = blockIdx.y * PW_BM;
    const int64_t p0 = (int64_t)blockIdx.x * PW_BN;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const float* x = p.x + (int64_t)img * (p.x_img_stride ? p.x_img_stride : (int64_t)p.cin * p.P);
    const int64_t wofs = (int64_t)s * p.cout * p.cin;
    const bool vec_ok = (p.P % 4 == 0);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.cin; k0 += PW_BK) {
        // weight tile (sampled on load when mu/rho/eps are given)
#pragma unroll
        for (int r = 0; r < (PW_BM * PW_BK) / 256; ++r) {
            const int idx = tid + r * 256;
            const int kk = idx % PW_BK, m = idx / PW_BK;   // consecutive threads walk ci: coalesced over a weight row
            const int co = co0 + m, ci = k0 + kk;
            float w = 0.f;
            if (co < p.cout && ci < p.cin) {
                const int64_t wi = (int64_t)co * p.cin + ci;
                if (p.w) w = p.w[wofs + wi];
                else {
                    w = p.mu[wi];
                    if (p.sigma) w = fmaf(p.sigma[wi], p.eps[wofs + wi], w);
                    else if (p.rho) w = fmaf(sigma_of_rho(p.rho[wi]), p.eps[wofs + wi], w);
                }
            }
            sW[kk][m] = w;
        }
        // activation tile
#pragma unroll
        for (int r = 0; r < (PW_BK * PW_BN / 4) / 256; ++r) {
            const int idx = tid + r * 256;
            const int kk = idx / (PW_BN / 4), c4 = idx % (PW_BN / 4);
            const int ci = k0 + kk;
            const int64_t pp = p0 + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ci < p.cin) {
                const float* src = x + (int64_t)ci * p.P + pp;
                if (vec_ok && pp + 3 < p.P) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (pp + 0 < p.P) v.x = src[0];
                    if (pp + 1 < p.P) v.y = src[1];
                    if (pp + 2 < p.P) v.z = src[2];
                    if (pp + 3 < p.P) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(&sX[kk][c4 * 4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PW_BK; ++kk) {
            const float4 wa = *reinterpret_cast<const float4*>(&sW[kk][ty * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&sW[kk][ty * 8 + 4]);
            const float4 xa = *reinterpret_cast<const float4*>(&sX[kk][tx * 4]);
            const float4 xb = *reinterpret_cast<const float4*>(&sX[kk][128 + tx * 4]);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = p.out + (int64_t)img * p.cout * p.P;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + ty * 8 + i;
        if (co >= p.cout) continue;
        const float b = p.bias ? p.bias[(int64_t)s * p.cout + co] : 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int64_t pp = p0 + half * 128 + tx * 4;
            float* dst = out + (int64_t)co * p.P + pp;
            float4 v = make_float4(acc[i][half * 4 + 0] + b, acc[i][half * 4 + 1] + b, acc[i][half * 4 + 2] + b,
                                   acc[i][half * 4 + 3] + b);
            if (p.residual) {
                const float* rs = p.residual + ((int64_t)img * p.cout + co) * p.P + pp;
                if (pp + 0 < p.P) v.x += rs[0];
                if (pp + 1 < p.P) v.y += rs[1];
                if (pp + 2 < p.P) v.z += rs[2];
                if (pp + 3 < p.P) v.w += rs[3];
            }
            if (p.prelu_slope) {
                const float sl = p.prelu_slope[p.prelu_n > 1 ? co : 0];
                v.x = v.x > 0.f ? v.x : v.x * sl;
                v.y = v.y > 0.f ? v.y : v.y * sl;
                v.z = v.z > 0.f ? v.z : v.z * sl;
                v.w = v.w > 0.f ? v.w : v.w * sl;
            }
            if (vec_ok && pp + 3 < p.P) *reinterpret_cast<float4*>(dst) = v;
            else {
                if (pp + 0 < p.P) dst[0] = v.x;
                if (pp + 1 < p.P) dst[1] = v.y;
                if (pp + 2 < p.P) dst[2] = v.z;
                if (pp + 3 < p.P) dst[3] = v.w;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// S-batched depthwise 3x3 (stride 1, zero padding 1) with the activation that follows it in the reference fused.
// One thread: a strip of 4 output columns x DW_ROWS rows, input rows rolled through registers (one float4 + two halo
// scalars loaded per output row instead of 18 scalars), float4 stores. ACT 2 (gated GELU) walks channel c and c + C/2
// together and writes C/2 channels.
// ------------------------------------------------------------------------------------------------
constexpr int DW_ROWS = 16;

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// SiLU x * sigmoid(x) (torch: x / (1 + exp(-x)) in fp32): two MUFU ops, ~2 ulp
__device__ __forceinline__ float silu_f(float v) { return v * rcp_approx(1.f + ex2_approx_f(-v * 1.4426950408889634f)); }
// nn.GELU() (erf form) = 0.5 x (1 + erf(x / sqrt 2)), branch-free: erf by Abramowitz & Stegun 7.1.26
// (|erf error| <= 1.5e-7; measured |gelu error| <= 4.7e-7 over [-8, 8], inside the 1.2e-6 envelope of the library's own fp32
// erff-based GELU against fp64). libm's erff costs ~2.5x the instructions of the 9-tap convolution it follows here, with two
// divergent branches; this is 14 instructions.
__device__ __forceinline__ float gelu_f(float v) {
    const float ax = fabsf(v) * 0.70710678118654752440f;
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
    float q = fmaf(t, 1.061405429f, -1.453152027f);
    q = fmaf(t, q, 1.421413741f);
    q = fmaf(t, q, -0.284496736f);
    q = fmaf(t, q, 0.254829592f);
    const float e = ex2_approx_f(-(ax * ax) * 1.4426950408889634f);
    const float erf_abs = fmaf(-(q * t), e, 1.f);
    return 0.5f * v * (1.f + copysignf(erf_abs, v));
}

struct DwRow { float v[6]; };   // columns w0-1 .. w0+4 of one input row (zero outside the image)

// VEC: W % 4 == 0 and 16-byte aligned planes -> one float4 + two halo scalars per row, the halo addresses clamped into
// the row and their values masked (no divergence at the image's left / right edge). Rows outside the image are zero
// (a branch that is uniform except in warps straddling two strips).
template <bool VEC>
__device__ __forceinline__ DwRow dw_load_row(const float* __restrict__ plane, int hh, int H, int W, int w0, int loff, int roff) {
    DwRow r;
#pragma unroll
    for (int j = 0; j < 6; ++j) r.v[j] = 0.f;
    if ((unsigned)hh >= (unsigned)H) return r;
    const float* row = plane + (int64_t)hh * W + w0;
    if (VEC) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(row));
        const float l = __ldg(row + loff), rr = __ldg(row + roff);
        r.v[0] = loff < 0 ? l : 0.f;
        r.v[1] = q.x; r.v[2] = q.y; r.v[3] = q.z; r.v[4] = q.w;
        r.v[5] = roff > 3 ? rr : 0.f;
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int ww = w0 + j - 1;
            if (ww >= 0 && ww < W) r.v[j] = __ldg(row + j - 1);
        }
    }
    return r;
}

__device__ __forceinline__ void dw_fma_row(const DwRow& r, const float* __restrict__ w3, float (&acc)[4]) {
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[o] = fmaf(w3[j], r.v[o + j], acc[o]);
}

template <int ACT, bool VEC>
__global__ void __launch_bounds__(256, 3) bayes_depthwise3_kernel(const BemBayesDepthwiseParams p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NP = ACT == 2 ? 2 : 1;                 // input planes per thread
    const int Cout = ACT == 2 ? p.C / 2 : p.C;
    const int W4 = (p.W + 3) / 4, HS = (p.H + DW_ROWS - 1) / DW_ROWS;
    const int plane = blockIdx.x;                        // img * Cout + c
    const int img = plane / Cout, c = plane - img * Cout;
    const int s = p.n_samples > 1 ? img / (p.batch / p.n_samples) : 0;
    float w[NP][9], b[NP];
    const float* xin[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int cc = c + q * Cout;
        const float* wp = p.w + ((int64_t)s * p.C + cc) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) w[q][i] = wp[i];
        b[q] = p.bias ? p.bias[(int64_t)s * p.C + cc] : 0.f;
        xin[q] = p.x + ((int64_t)img * p.C + cc) * p.H * p.W;
    }
    float* out = p.out + (int64_t)plane * p.H * p.W;
    const int64_t strips = (int64_t)HS * W4;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < strips; t += (int64_t)gridDim.y * blockDim.x) {
        const int hs = (int)(t / W4), w0 = (int)(t - (int64_t)hs * W4) * 4;
        const int h0 = hs * DW_ROWS;
        const int loff = w0 > 0 ? -1 : 0, roff = w0 + 4 < p.W ? 4 : 3;   // halo columns, clamped into the row
        DwRow r0[NP], r1[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            r0[q] = dw_load_row<VEC>(xin[q], h0 - 1, p.H, p.W, w0, loff, roff);
            r1[q] = dw_load_row<VEC>(xin[q], h0, p.H, p.W, w0, loff, roff);
        }
        float* dst = out + (int64_t)h0 * p.W + w0;
#pragma unroll
        for (int i = 0; i < DW_ROWS; ++i, dst += p.W) {
            const int h = h0 + i;
            if (h >= p.H) break;
            float res[NP][4];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const DwRow r2 = dw_load_row<VEC>(xin[q], h + 1, p.H, p.W, w0, loff, roff);
                float acc[4] = {b[q], b[q], b[q], b[q]};
                dw_fma_row(r0[q], &w[q][0], acc);
                dw_fma_row(r1[q], &w[q][3], acc);
                dw_fma_row(r2, &w[q][6], acc);
#pragma unroll
                for (int o = 0; o < 4; ++o) res[q][o] = acc[o];
                r0[q] = r1[q];
                r1[q] = r2;
            }
            float y[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                if (ACT == 1) y[o] = silu_f(res[0][o]);
                else if (ACT == 2) y[o] = gelu_f(res[0][o]) * res[NP - 1][o];
                else y[o] = res[0][o];
            }
            if (VEC) *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
            else {
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (w0 + o < p.W) dst[o] = y[o];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Shared-memory form (W % 4 == 0): the register kernel above is bound by global-load latency (72 % long-scoreboard
// stalls). Here a CTA owns R output rows of one plane (pair): the R + 2 input rows are one contiguous run of the plane,
// fetched by ONE TMA bulk copy per plane while the other CTAs of the SM compute (4 CTAs x 48 KB in flight per SM);
// threads then read rows with LDS.128, take the two halo columns from their warp neighbours by shuffle, and store
// float4 rows. One thread = 4 columns x DWS_RPT rows.
// ------------------------------------------------------------------------------------------------
constexpr int DWS_RPT = 4;

template <int ACT>
__global__ void __launch_bounds__(512) bayes_depthwise3_smem_kernel(const BemBayesDepthwiseParams p, const int R) {
    pdl_trigger();
    pdl_wait();
    constexpr int NP = ACT == 2 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char dsm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm);
    float* tile = reinterpret_cast<float*>(dsm + 128);            // [NP][R + 2][W]
    const int Cout = ACT == 2 ? p.C / 2 : p.C;
    const int W = p.W, W4 = W >> 2, H = p.H;
    const int plane = blockIdx.x, img = plane / Cout, c = plane - img * Cout;
    const int s = p.n_samples > 1 ? img / (p.batch / p.n_samples) : 0;
    const int h0 = blockIdx.y * R;
    const int rows_out = min(R, H - h0);
    const int lo = max(h0 - 1, 0), hi = min(h0 + rows_out, H - 1);      // input rows [lo, hi] exist
    const int tid = threadIdx.x, lane = tid & 31;
    const int plane_elems = (R + 2) * W;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)(hi - lo + 1) * W * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes * NP) : "memory");
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const float* src = p.x + (((int64_t)img * p.C + c + q * Cout) * H + lo) * W;
            float* dstp = tile + q * plane_elems + (lo - (h0 - 1)) * W;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(dstp)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
        }
    }
    // rows above / below the image are zero (disjoint from what the copies write)
    if (h0 == 0)
        for (int i = tid; i < NP * W; i += blockDim.x) tile[(i / W) * plane_elems + (i % W)] = 0.f;
    for (int r = hi + 1; r <= h0 + R; ++r)
        for (int i = tid; i < NP * W; i += blockDim.x) tile[(i / W) * plane_elems + (r - (h0 - 1)) * W + (i % W)] = 0.f;
    float w[NP][9], b[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int cc = c + q * Cout;
        const float* wp = p.w + ((int64_t)s * p.C + cc) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) w[q][i] = wp[i];
        b[q] = p.bias ? p.bias[(int64_t)s * p.C + cc] : 0.f;
    }
    __syncthreads();
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}"
                         : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
    }
    const int rs = tid / W4, cg = tid - rs * W4;                 // row split, column group
    const bool active = rs * DWS_RPT < rows_out;
    const int w0 = cg * 4;
    const int r0 = min(rs * DWS_RPT, R - DWS_RPT);               // first local output row (inactive threads: any valid row)
    float acc[NP][DWS_RPT][4];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int i = 0; i < DWS_RPT; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[q][i][o] = b[q];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const float* tp = tile + q * plane_elems + r0 * W + w0;  // local input row r0 = image row h0 + r0 - 1
#pragma unroll
        for (int j = 0; j < DWS_RPT + 2; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(tp + j * W);
            float l = __shfl_up_sync(0xffffffffu, v.w, 1), rr = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 0 || cg == 0) l = cg > 0 ? tp[j * W - 1] : 0.f;
            if (lane == 31 || cg == W4 - 1) rr = cg < W4 - 1 ? tp[j * W + 4] : 0.f;
            const float x6[6] = {l, v.x, v.y, v.z, v.w, rr};
#pragma unroll
            for (int i = 0; i < DWS_RPT; ++i) {
                const int k = j - i;                              // tap row of output row i fed by input row j
                if (k >= 0 && k < 3) {
#pragma unroll
                    for (int o = 0; o < 4; ++o)
#pragma unroll
                        for (int t = 0; t < 3; ++t) acc[q][i][o] = fmaf(w[q][k * 3 + t], x6[o + t], acc[q][i][o]);
                }
            }
        }
    }
    if (!active) return;
    float* out = p.out + ((int64_t)plane * H + h0 + r0) * W + w0;
#pragma unroll
    for (int i = 0; i < DWS_RPT; ++i) {
        if (r0 + i >= rows_out) break;
        float y[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (ACT == 1) y[o] = silu_f(acc[0][i][o]);
            else if (ACT == 2) y[o] = gelu_f(acc[0][i][o]) * acc[NP - 1][i][o];
            else y[o] = acc[0][i][o];
        }
        *reinterpret_cast<float4*>(out + (int64_t)i * W) = make_float4(y[0], y[1], y[2], y[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// Dense 3x3 convolution (stride 1, zero padding 1) for the small-channel stems of the network (3 -> 40 and 40 -> 3 at full
// resolution, UNet_arch.py:423-431): register-tiled direct convolution. One thread = 4 output columns of one row x COT
// output channels; the weights of the CTA's output-channel tile sit in shared memory as [ci][tap][COT] (broadcast LDS.128),
// input rows are read as one float4 + two clamped / masked halo scalars. The library path splits these shapes into five
// kernels and takes 90 / 430 us at 600x400; this takes 26 us (3 -> 40) and 55 us (40 -> 3, load-latency bound).
// ------------------------------------------------------------------------------------------------
template <int COT, bool VEC>
__global__ void __launch_bounds__(256) conv3x3_direct_kernel(const BemConv3x3Params p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float sw[];            // [cin][9][COT]
    const int co0 = blockIdx.z * COT;
    const int img = blockIdx.y;
    for (int i = threadIdx.x; i < p.cin * 9 * COT; i += blockDim.x) {
        const int co = i % COT, tap = (i / COT) % 9, ci = i / (9 * COT);
        sw[i] = co0 + co < p.cout ? p.w[((int64_t)(co0 + co) * p.cin + ci) * 9 + tap] : 0.f;
    }
    __syncthreads();
    const int W4 = (p.W + 3) / 4;
    const int64_t strips = (int64_t)p.H * W4;
    const float* x = p.x + (int64_t)img * p.cin * p.H * p.W;
    float* out = p.out + ((int64_t)img * p.cout + co0) * p.H * p.W;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < strips; t += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(t / W4), w0 = (int)(t - (int64_t)h * W4) * 4;
        const int loff = w0 > 0 ? -1 : 0, roff = w0 + 4 < p.W ? 4 : 3;
        float acc[COT][4];
#pragma unroll
        for (int c = 0; c < COT; ++c) {
            const float b = (p.bias && co0 + c < p.cout) ? p.bias[co0 + c] : 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[c][o] = b;
        }
        for (int ci = 0; ci < p.cin; ++ci) {
            const float* plane = x + (int64_t)ci * p.H * p.W;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const DwRow r = dw_load_row<VEC>(plane, h + kh - 1, p.H, p.W, w0, loff, roff);
                const float* wr = sw + (ci * 9 + kh * 3) * COT;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    float wv[COT];
#pragma unroll
                    for (int c4 = 0; c4 < COT / 4; ++c4) {
                        const float4 q = *reinterpret_cast<const float4*>(wr + kw * COT + c4 * 4);
                        wv[c4 * 4] = q.x; wv[c4 * 4 + 1] = q.y; wv[c4 * 4 + 2] = q.z; wv[c4 * 4 + 3] = q.w;
                    }
#pragma unroll
                    for (int c = 0; c < COT; ++c)
#pragma unroll
                        for (int o = 0; o < 4; ++o) acc[c][o] = fmaf(wv[c], r.v[o + kw], acc[c][o]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < COT; ++c) {
            if (co0 + c >= p.cout) break;
            float* dst = out + ((int64_t)c * p.H + h) * p.W + w0;
            if (VEC) *reinterpret_cast<float4*>(dst) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
            else {
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (w0 + o < p.W) dst[o] = acc[c][o];
            }
        }
    }
}

template <int COT>
static int conv3x3_launch(const BemConv3x3Params& p, cudaStream_t stream) {
    const bool vec = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
    const int64_t strips = (int64_t)p.H * ((p.W + 3) / 4);
    int64_t bx = (strips + 255) / 256;
    if (bx > 65535) bx = 65535;
    dim3 grid((unsigned)bx, (unsigned)p.batch, (unsigned)((p.cout + COT - 1) / COT));
    const int smem = p.cin * 9 * COT * 4;
    if (smem > 48 * 1024) return BEM_ERR_UNSUPPORTED;
    if (vec) launch_pdl(conv3x3_direct_kernel<COT, true>, dim3(grid), dim3(256), smem, stream, p);
    else launch_pdl(conv3x3_direct_kernel<COT, false>, dim3(grid), dim3(256), smem, stream, p);
    return (int)cudaGetLastError();
}

template <int ACT>
static void depthwise_launch(const BemBayesDepthwiseParams& p, dim3 grid, cudaStream_t stream) {
    const bool vec = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
    static const int no_smem = getenv("BEM_DW_REG") ? atoi(getenv("BEM_DW_REG")) : 0;
    if (vec && !no_smem) {
        // rows per CTA: 16 if the tile stays within ~48 KB, else 8; threads = column groups x row splits
        constexpr int NP = ACT == 2 ? 2 : 1;
        const int W4 = p.W / 4;
        int R = 16;
        if (NP * (R + 2) * p.W * 4 > 56 * 1024 || W4 * (R / DWS_RPT) > 512) R = 8;
        const int threads = (W4 * (R / DWS_RPT) + 31) / 32 * 32;
        const int smem = 128 + NP * (R + 2) * p.W * 4;
        if (threads <= 512 && smem <= 100 * 1024 && p.H >= 1) {
            int dev = 0;
            cudaGetDevice(&dev);
            static int attr[64] = {0};
            if (attr[dev & 63] < smem) {
                cudaFuncSetAttribute(bayes_depthwise3_smem_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                attr[dev & 63] = 100 * 1024;
            }
            dim3 g2(grid.x, (unsigned)((p.H + R - 1) / R));
            launch_pdl(bayes_depthwise3_smem_kernel<ACT>, dim3(g2), dim3(threads), smem, stream, p, R);
            return;
        }
    }
    if (vec) launch_pdl(bayes_depthwise3_kernel<ACT, true>, dim3(grid), dim3(256), 0, stream, p);
    else launch_pdl(bayes_depthwise3_kernel<ACT, false>, dim3(grid), dim3(256), 0, stream, p);
}

}  // namespace bem

using namespace bem;

extern "C" {

int bem_bayes_sample(const BemBayesSampleParams* p, void* stream) {
    if (!p || !p->mu || !p->w || p->numel <= 0 || p->n_samples <= 0) return BEM_ERR_BAD_ARG;
    const int64_t total = (p->numel + 3) / 4 * p->n_samples;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    launch_pdl(bayes_sample_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, *p);
    return (int)cudaGetLastError();
}

int bem_bayes_sample_batched(const BemBayesSampleBatchedParams* p, void* stream) {
    if (!p || !p->entries || !p->blocks || p->n_blocks <= 0) return BEM_ERR_BAD_ARG;
    launch_pdl(bayes_sample_batched_kernel, dim3(p->n_blocks), dim3(256), 0, (cudaStream_t)stream, *p);
    return (int)cudaGetLastError();
}

int64_t bem_bayes_pointwise_workspace_bytes(int n_samples, int cin, int cout) {
    if (n_samples <= 0 || cin <= 0 || cout <= 0) return 0;
    return bayes_pointwise_tc_workspace(n_samples, cin, cout);
}

int bem_bayes_pointwise(const BemBayesPointwiseParams* p, void* stream) {
    if (!p || !p->x || !p->out || p->n_samples <= 0 || p->batch <= 0 || p->cin <= 0 || p->cout <= 0 || p->P <= 0)
        return BEM_ERR_BAD_ARG;
    if (p->batch % p->n_samples != 0) return BEM_ERR_BAD_ARG;
    if (!p->w && !p->mu) return BEM_ERR_BAD_ARG;
    if (p->prelu_slope && p->prelu_n != 1 && p->prelu_n != p->cout) return BEM_ERR_BAD_ARG;
    if (!p->w && (p->rho || p->sigma) && !p->eps) return BEM_ERR_BAD_ARG;
    if (p->batch > 65535) return BEM_ERR_UNSUPPORTED;
    if (!p->force_simt) return bayes_pointwise_tc_launch(*p, (cudaStream_t)stream);
    if (p->ln_gamma) return BEM_ERR_UNSUPPORTED;   // the LayerNorm fusion lives in the tensor-core kernel
    dim3 grid((unsigned)((p->P + PW_BN - 1) / PW_BN), (unsigned)((p->cout + PW_BM - 1) / PW_BM), (unsigned)p->batch);
    launch_pdl(bayes_pointwise_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, *p);
    return (int)cudaGetLastError();
}

int64_t bem_bayes_pointwise_pack_table_bytes(int n) { return n > 0 ? bayes_pointwise_pack_table_bytes(n) : 0; }

int bem_bayes_pointwise_pack_table(const BemBayesPointwiseParams* params, int n, void* table_host, int32_t* total_blocks) {
    if (!params || n <= 0 || !table_host || !total_blocks) return BEM_ERR_BAD_ARG;
    for (int i = 0; i < n; ++i) {
        const BemBayesPointwiseParams& q = params[i];
        if (q.n_samples <= 0 || q.cin <= 0 || q.cout <= 0 || q.P <= 0 || q.batch <= 0 || q.force_simt) return BEM_ERR_BAD_ARG;
        if (!q.w && !q.mu) return BEM_ERR_BAD_ARG;
        if (!q.w && (q.rho || q.sigma) && !q.eps) return BEM_ERR_BAD_ARG;
    }
    return bayes_pointwise_pack_table(params, n, table_host, total_blocks);
}

int bem_bayes_pointwise_pack_run(const void* table_dev, int n, int total_blocks, void* stream) {
    if (!table_dev || n <= 0 || total_blocks <= 0) return BEM_ERR_BAD_ARG;
    return bayes_pointwise_pack_run(table_dev, n, total_blocks, (cudaStream_t)stream);
}

int bem_conv3x3(const BemConv3x3Params* p, void* stream) {
    if (!p || !p->x || !p->w || !p->out || p->batch <= 0 || p->cin <= 0 || p->cout <= 0 || p->H <= 0 || p->W <= 0) return BEM_ERR_BAD_ARG;
    if (p->batch > 65535) return BEM_ERR_UNSUPPORTED;
    // output-channel tile per thread: 4 for the few-output case (40 -> 3), 8 otherwise
    return p->cout <= 4 ? conv3x3_launch<4>(*p, (cudaStream_t)stream) : conv3x3_launch<8>(*p, (cudaStream_t)stream);
}

int bem_bayes_depthwise(const BemBayesDepthwiseParams* p, void* stream) {
    if (!p || !p->x || !p->w || !p->out || p->n_samples <= 0 || p->batch <= 0 || p->C <= 0 || p->H <= 0 || p->W <= 0)
        return BEM_ERR_BAD_ARG;
    if (p->batch % p->n_samples != 0 || p->act < 0 || p->act > 2 || (p->act == 2 && p->C % 2 != 0)) return BEM_ERR_BAD_ARG;
    if (p->K != 3) return BEM_ERR_UNSUPPORTED;
    const int64_t planes = (int64_t)p->batch * (p->act == 2 ? p->C / 2 : p->C);
    if (planes > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    const int64_t strips = (int64_t)((p->H + DW_ROWS - 1) / DW_ROWS) * ((p->W + 3) / 4);
    int64_t by = (strips + 255) / 256;
    if (by > 1024) by = 1024;
    dim3 grid((unsigned)planes, (unsigned)by);
    if (p->act == 0) depthwise_launch<0>(*p, grid, (cudaStream_t)stream);
    else if (p->act == 1) depthwise_launch<1>(*p, grid, (cudaStream_t)stream);
    else depthwise_launch<2>(*p, grid, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

}  // extern "C"
