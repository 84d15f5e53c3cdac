// bayes_tc.cu — S-batched 1x1 convolution with per-sample (Bayesian) weights on the 5th-generation tensor cores.
//
// Replaces, for the Bayesian 1x1 layers (Linear2dReparameterization / Conv2dReparameterization with a 1x1 kernel,
// basicsr/bayesian/linear.py:82-90, conv.py:106-114), the eager chain `sigma = log1p(exp(rho)); w = mu + sigma * eps;
// F.conv2d(x, w, b)` and — when the layer is preceded by a LayerNorm2d, as every Bayesian 1x1 of a VSSBlock is
// (vmamba.py:1319-1334, :696-715) — that normalisation as well.
//
//   D[p][co] = sum_ci xhat[ci][p] * W[s][co][ci]        M = 128 pixels (TMEM lanes), N = output-channel tile, K = ci
//
//   * tcgen05.mma.cta_group::1.kind::tf32, both operands K-major in the no-swizzle canonical layout (8-row x 16-byte
//     core matrices), accumulators in TMEM (256 columns per CTA, two CTAs per SM).
//   * fp32 parity: every operand is split into a tf32 "hi" part and the fp32 remainder "lo"; three MMAs per K step
//     (hi*hi + hi*lo + lo*hi) give ~2^-21 relative accuracy, i.e. the 1e-5 tier of the parity tests. The contraction is
//     HBM-bound at these channel counts (arithmetic intensity 20-140 FLOP/B), so the 3x tensor work is free.
//   * the operand staging is where the fusion happens: the 128 threads of a CTA read x coalesced along pixels, apply
//     the (optional) LayerNorm with per-pixel statistics, split, and store K-major. The weights are sampled
//     (mu + sigma * eps), split and laid out once per launch by a tiny pack kernel, and reach shared memory with one
//     TMA bulk copy per K chunk; a sampled fp32 weight tensor is never materialised.
//   * epilogue: tcgen05.ld (32 lanes x 32 bit x 16 columns) -> + bias -> 128-byte coalesced stores along pixels.
#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

constexpr int TC_M = 128;    // pixels per CTA tile (TMEM lanes)
constexpr int TC_KC = 16;    // K elements staged per pipeline step (two MMA K-steps of 8)
constexpr int TC_NMAX = 256; // output channels per CTA tile (TMEM columns)

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major, no-swizzle shared-memory descriptor (cute::UMMA::SmemDescriptor, version 1):
// start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// canonical K-major layout of one [rows x TC_KC] fp32 tile: core matrix = 8 rows x 16 bytes, K-adjacent core matrices
// 128 B apart (LBO), 8-row groups TC_KC/4 * 128 B apart (SBO)
constexpr uint32_t TC_LBO = 128;
constexpr uint32_t TC_SBO = (TC_KC / 4) * 128;
__device__ __forceinline__ uint32_t tile_off(int row, int k4) { return (uint32_t)((row >> 3) * TC_SBO + k4 * TC_LBO + (row & 7) * 16); }

// ------------------------------------------------------------------------------------------------
// weight pack: sample (w = mu + sigma * eps), split into tf32 hi / fp32 remainder lo, and lay the tiles out exactly as
// the MMA wants them in shared memory, so the GEMM fetches its B operand with one TMA bulk copy per K chunk.
// pack[((s * ntiles + tile) * nk + kc)] = [hi tile NT x KC | lo tile NT x KC], canonical K-major layout (tile_off)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bayes_weight_pack_kernel(const BemBayesPointwiseParams p, const int NT, const int ntiles,
                                                              const int nk, float* __restrict__ pack) {
    const int blk = blockIdx.x;             // (s, tile, kc)
    const int kc = blk % nk;
    const int tile = (blk / nk) % ntiles;
    const int s = blk / (nk * ntiles);
    const int n0 = tile * NT, k0 = kc * TC_KC;
    const int64_t wofs = (int64_t)s * p.cout * p.cin;
    unsigned char* hi_t = reinterpret_cast<unsigned char*>(pack + (int64_t)blk * 2 * NT * TC_KC);
    unsigned char* lo_t = hi_t + (size_t)NT * TC_KC * 4;
    for (int idx = threadIdx.x; idx < NT * (TC_KC / 4); idx += blockDim.x) {
        const int n = idx / (TC_KC / 4), k4 = idx - n * (TC_KC / 4);
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ci = k0 + k4 * 4 + e, co = n0 + n;
            float w = 0.f;
            if (co < p.cout && ci < p.cin) {
                const int64_t wi = (int64_t)co * p.cin + ci;
                if (p.w) w = p.w[wofs + wi];
                else {
                    w = p.mu[wi];
                    if (p.sigma) w = fmaf(p.sigma[wi], p.eps[wofs + wi], w);
                    else if (p.rho) w = fmaf(log1pf(expf(p.rho[wi])), p.eps[wofs + wi], w);
                }
            }
            hi[e] = tf32_hi(w);
            lo[e] = w - hi[e];
        }
        const uint32_t off = tile_off(n, k4);
        *reinterpret_cast<float4*>(hi_t + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(lo_t + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

__global__ void __launch_bounds__(128, 4) bayes_pointwise_tc_kernel(const BemBayesPointwiseParams p, const int NT, const int ntiles,
                                                                  const float* __restrict__ pack, const uint32_t tmem_cols) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // stage s: [A hi | A lo | B hi | B lo]; A tiles 128 x KC (staged by the threads), B tiles NT x KC (TMA from `pack`)
    const uint32_t a_bytes = TC_M * TC_KC * 4;
    const uint32_t b_bytes = (uint32_t)NT * TC_KC * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    unsigned char* tail = smem + 2 * stage_bytes;
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(tail);        // [2] MMAs reading a stage have completed
    uint64_t* b_full = mma_done + 2;                                // [2] B tiles of a stage have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 32);
    float* s_gamma = reinterpret_cast<float*>(tail + 48);           // [cin] LayerNorm weight / bias (optional)
    float* s_beta = s_gamma + p.cin;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int n0 = tile * NT;                                       // first output channel of this tile
    const int64_t p0 = (int64_t)blockIdx.y * TC_M;
    const int img = blockIdx.z;
    const int s_idx = p.n_samples > 1 ? (p.sample_interleave ? img % p.n_samples : img / (p.batch / p.n_samples)) : 0;
    const int nvalid = min(NT, p.cout - n0);
    const float* x = p.x + (int64_t)img * (p.x_img_stride ? p.x_img_stride : (int64_t)p.cin * p.P);
    const int64_t pix = p0 + tid;
    const bool pvalid = pix < p.P;
    const int nk = (p.cin + TC_KC - 1) / TC_KC;
    const float* bsrc = pack + ((int64_t)(s_idx * ntiles + tile) * nk) * 2 * NT * TC_KC;

    if (tid == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        mbar_init(&b_full[0], 1);
        mbar_init(&b_full[1], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    const bool ln = p.ln_gamma != nullptr;
    if (ln) {
        for (int i = tid; i < p.cin; i += 128) {
            s_gamma[i] = p.ln_gamma[i];
            s_beta[i] = p.ln_beta ? p.ln_beta[i] : 0.f;
        }
    }
    // LayerNorm statistics of this thread's pixel over the input channels: one pass over sums shifted by the first
    // channel (well conditioned), all loads independent
    float mean = 0.f, rstd = 1.f;
    if (ln && pvalid) {
        const float x0 = x[pix];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
        for (int ci = 1; ci < p.cin; ++ci) {
            const float dlt = x[(int64_t)ci * p.P + pix] - x0;
            s1 += dlt;
            s2 = fmaf(dlt, dlt, s2);
        }
        const float inv = 1.f / (float)p.cin;
        const float m1 = s1 * inv;
        mean = x0 + m1;
        rstd = rsqrtf(fmaxf(s2 * inv - m1 * m1, 0.f) + p.ln_eps);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // instruction descriptor: D fp32, A/B tf32, K-major both, N = NT, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    for (int kc = 0; kc < nk; ++kc) {
        const int st = kc & 1;
        const uint32_t use = kc >> 1;
        if (kc >= 2) mbar_wait(&mma_done[st], (use - 1) & 1, nullptr);   // MMAs that read this stage are done
        unsigned char* sA_hi = smem + (size_t)st * stage_bytes;
        unsigned char* sA_lo = sA_hi + a_bytes;
        unsigned char* sB_hi = sA_lo + a_bytes;   // lo tile follows contiguously, as in `pack`
        if (tid == 0) {
            mbar_arrive_expect_tx(&b_full[st], 2 * b_bytes);
            bulk_g2s(sB_hi, bsrc + (int64_t)kc * 2 * NT * TC_KC, 2 * b_bytes, &b_full[st]);
        }
        const int k0 = kc * TC_KC;
        // ---- activation tile: row = this thread's pixel, 16 channels (loads first, then the arithmetic) ----
        float xv[TC_KC];
#pragma unroll
        for (int e = 0; e < TC_KC; ++e) {
            const int ci = k0 + e;
            xv[e] = (pvalid && ci < p.cin) ? x[(int64_t)ci * p.P + pix] : 0.f;
        }
#pragma unroll
        for (int k4 = 0; k4 < TC_KC / 4; ++k4) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ci = k0 + k4 * 4 + e;
                float v = xv[k4 * 4 + e];
                if (ln && pvalid && ci < p.cin) v = fmaf((v - mean) * rstd, s_gamma[ci], s_beta[ci]);
                hi[e] = tf32_hi(v);
                lo[e] = v - hi[e];
            }
            const uint32_t off = tile_off(tid, k4);
            *reinterpret_cast<float4*>(sA_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(sA_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            mbar_wait(&b_full[st], use & 1, nullptr);
            tc_fence_after();
            const uint32_t aH = smem_u32(sA_hi), aL = smem_u32(sA_lo), bH = smem_u32(sB_hi), bL = bH + b_bytes;
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {
                const uint32_t adv = ks * 2 * TC_LBO;   // 8 tf32 = two 16-byte K chunks per MMA
                const uint64_t dAh = make_desc(aH + adv, TC_LBO, TC_SBO), dAl = make_desc(aL + adv, TC_LBO, TC_SBO);
                const uint64_t dBh = make_desc(bH + adv, TC_LBO, TC_SBO), dBl = make_desc(bL + adv, TC_LBO, TC_SBO);
                umma_tf32(tmem, dAh, dBh, idesc, (kc | ks) != 0);
                umma_tf32(tmem, dAh, dBl, idesc, 1);
                umma_tf32(tmem, dAl, dBh, idesc, 1);
            }
            umma_commit(&mma_done[st]);   // implies tcgen05.fence::before_thread_sync
        }
    }
    // all MMAs complete when the last commit has arrived (commits complete in order)
    {
        const int last = nk - 1;
        mbar_wait(&mma_done[last & 1], (last >> 1) & 1, nullptr);
        tc_fence_after();
    }
    // ---- epilogue: TMEM -> registers -> + bias -> global, 16 output channels at a time ----
    float* out = p.out + (int64_t)img * p.cout * p.P;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < nvalid; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_base + (uint32_t)c0, v);
        if (pvalid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int co = n0 + c0 + i;
                if (c0 + i < nvalid) {
                    const float b = p.bias ? p.bias[(int64_t)s_idx * p.cout + co] : 0.f;
                    out[(int64_t)co * p.P + pix] = v[i] + b;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

static void tc_tiling(int cin, int cout, int& ntiles, int& NT, int& nk) {
    // output-channel tiling: as few tiles as possible, each a multiple of 16 and at most 256 channels
    ntiles = (cout + TC_NMAX - 1) / TC_NMAX;
    NT = ((cout + ntiles - 1) / ntiles + 15) / 16 * 16;
    if (NT < 16) NT = 16;
    nk = (cin + TC_KC - 1) / TC_KC;
}

int64_t bayes_pointwise_tc_workspace(int n_samples, int cin, int cout) {
    int ntiles, NT, nk;
    tc_tiling(cin, cout, ntiles, NT, nk);
    return (int64_t)n_samples * ntiles * nk * 2 * NT * TC_KC * (int64_t)sizeof(float);
}

int bayes_pointwise_tc_launch(const BemBayesPointwiseParams& p, cudaStream_t stream) {
    int ntiles, NT, nk;
    tc_tiling(p.cin, p.cout, ntiles, NT, nk);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NT) tmem_cols <<= 1;
    const int smem_bytes = 2 * (2 * TC_M * TC_KC * 4 + 2 * NT * TC_KC * 4) + 48 + 2 * p.cin * 4 + 64;
    if (smem_bytes > 227 * 1024) return BEM_ERR_UNSUPPORTED;
    const int64_t need = bayes_pointwise_tc_workspace(p.n_samples, p.cin, p.cout);
    if (!p.workspace || p.workspace_bytes < need || (reinterpret_cast<uintptr_t>(p.workspace) & 15)) return BEM_ERR_WORKSPACE;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    static int attr_set[64] = {0};
    if (attr_set[dev] < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(bayes_pointwise_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = smem_bytes;
    }
    const int64_t ptiles = (p.P + TC_M - 1) / TC_M;
    if (ptiles > 65535 || p.batch > 65535) return BEM_ERR_UNSUPPORTED;
    float* pack = reinterpret_cast<float*>(p.workspace);
    bayes_weight_pack_kernel<<<p.n_samples * ntiles * nk, 256, 0, stream>>>(p, NT, ntiles, nk, pack);
    dim3 grid((unsigned)ntiles, (unsigned)ptiles, (unsigned)p.batch);
    bayes_pointwise_tc_kernel<<<grid, 128, smem_bytes, stream>>>(p, NT, ntiles, pack, tmem_cols);
    return (int)cudaGetLastError();
}

}  // namespace bem
