// bayes_tc.cu — S-batched 1x1 convolution with per-sample (Bayesian) weights on the 5th-generation tensor cores.
//
// Replaces, for the Bayesian 1x1 layers (Linear2dReparameterization / Conv2dReparameterization with a 1x1 kernel,
// basicsr/bayesian/linear.py:82-90, conv.py:106-114), the eager chain `sigma = log1p(exp(rho)); w = mu + sigma * eps;
// F.conv2d(x, w, b)` and — when the layer is preceded by a LayerNorm2d, as every Bayesian 1x1 of a VSSBlock is
// (vmamba.py:1319-1334, :696-715) — that normalisation as well.
//
//   D[p][co] = sum_ci xhat[ci][p] * W[s][co][ci]        M = 128 pixels (TMEM lanes), N = output-channel tile, K = ci
//
// Common to both kernels of this file:
//   * tcgen05.mma.cta_group::1.kind::tf32, accumulators in TMEM; the B operand (weights) K-major in the no-swizzle
//     canonical shared-memory layout (8-row x 16-byte core matrices), written once per launch by a pack kernel.
//   * fp32 parity: every operand is split into a tf32 "hi" part and the fp32 remainder "lo"; three MMAs per K step
//     (hi*hi + hi*lo + lo*hi) give ~2^-21 relative accuracy, i.e. the 1e-5 tier of the parity tests. The contraction is
//     HBM-bound at these channel counts (arithmetic intensity 20-140 FLOP/B), so the 3x tensor work is free.
//   * the sampled weight (mu + sigma * eps) is formed by the pack kernel; a sampled fp32 weight tensor is never materialised.
//   * epilogue: tcgen05.ld (32 lanes x 32 bit x 16 columns) -> per-channel affine (+ skip connection) -> 128-byte
//     coalesced stores along pixels.
// bayes_pointwise_tc3_kernel (the default; "persistent, warp-specialised form" below): cp.async activation ring, A operand
//   in TMEM, LayerNorm folded into weights + epilogue, resident or streamed weight tiles.
// bayes_pointwise_tc_kernel: one CTA per (channel tile, pixel tile), A staged through shared memory by the CTA's threads
//   with the LayerNorm applied on the way — kept for inputs whose pixel rows are not 16-byte aligned (cp.async cannot
//   move them) and for workloads whose epilogue vectors do not fit the persistent kernel's shared memory.
#include <algorithm>
#include <cstdlib>

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
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 8 consecutive TMEM columns of this thread's lane <- registers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&va)[16], float (&vb)[16]) {
    uint32_t r[16], q[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(tb)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        va[i] = __uint_as_float(r[i]);
        vb[i] = __uint_as_float(q[i]);
    }
}
// K-major, no-swizzle shared-memory descriptor (cute::UMMA::SmemDescriptor, version 1):
// start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
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
__device__ __forceinline__ float sampled_weight(const BemBayesPointwiseParams& p, int64_t wofs, int64_t wi) {
    if (p.w) return p.w[wofs + wi];
    float w = p.mu[wi];
    if (p.sigma) w = fmaf(p.sigma[wi], p.eps[wofs + wi], w);
    else if (p.rho) w = fmaf(log1pf(expf(p.rho[wi])), p.eps[wofs + wi], w);
    return w;
}

// `fold_ln` (persistent kernel): the LayerNorm weight is folded into the packed tiles, W' = gamma * W, and the blocks
// past the tile blocks write, per (sample, output channel n), vec = ( s_n = sum_ci W'[n][ci], t_n = sum_ci beta[ci] *
// W[n][ci] + bias[n] ), so that LN(x) . W = rstd * (x . W' - mean * s) + t is finished in the GEMM epilogue.
__device__ __forceinline__ void weight_pack_block(const BemBayesPointwiseParams& p, const int NT, const int ntiles, const int nk,
                                                  float* __restrict__ pack, const int fold_ln, float* __restrict__ vec, const int block) {
    const int tile_blocks = p.n_samples * ntiles * nk;
    if (block >= tile_blocks) {
        const int cblocks = (p.cout + 7) / 8;
        const int vb = block - tile_blocks;
        const int s = vb / cblocks, co = (vb - s * cblocks) * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        if (co >= p.cout) return;
        const int64_t wofs = (int64_t)s * p.cout * p.cin;
        float ss = 0.f, tt = 0.f;
        if (p.ln_gamma) {
            for (int ci = lane; ci < p.cin; ci += 32) {
                const float w = sampled_weight(p, wofs, (int64_t)co * p.cin + ci);
                ss = fmaf(w, p.ln_gamma[ci], ss);
                if (p.ln_beta) tt = fmaf(w, p.ln_beta[ci], tt);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
                tt += __shfl_xor_sync(0xffffffffu, tt, o);
            }
        }
        if (lane == 0) {
            if (p.bias) tt += p.bias[(int64_t)s * p.cout + co];
            float* v = vec + ((int64_t)s * ntiles + co / NT) * 2 * NT;
            v[2 * (co % NT)] = ss;
            v[2 * (co % NT) + 1] = tt;
        }
        return;
    }
    const int blk = block;                  // (s, tile, kc)
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
                w = sampled_weight(p, wofs, (int64_t)co * p.cin + ci);
                if (fold_ln && p.ln_gamma) w *= p.ln_gamma[ci];
            }
            hi[e] = tf32_hi(w);
            lo[e] = w - hi[e];
        }
        const uint32_t off = tile_off(n, k4);
        *reinterpret_cast<float4*>(hi_t + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(lo_t + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

__global__ void __launch_bounds__(256) bayes_weight_pack_kernel(const BemBayesPointwiseParams p, const int NT, const int ntiles,
                                                              const int nk, float* __restrict__ pack, const int fold_ln,
                                                              float* __restrict__ vec) {
    pdl_trigger();
    pdl_wait();
    weight_pack_block(p, NT, ntiles, nk, pack, fold_ln, vec, (int)blockIdx.x);
}

// The pack step of MANY layers in one launch (bem_bayes_pointwise_pack_run): a Monte-Carlo forward re-draws every Bayesian
// weight first (bem_bayes_sample_batched), so all its 1x1 layers can be packed right after, and each layer's own call then
// runs with `prepacked` — one launch per sample instead of one per layer. Block b belongs to the last entry whose first
// block is <= b.
struct PackEntry {
    BemBayesPointwiseParams p;
    float* pack;
    float* vec;
    int32_t NT, ntiles, nk, fold_ln, block0, nblocks;
};
static_assert(sizeof(PackEntry) % 4 == 0 && sizeof(PackEntry) <= 1024, "PackEntry is staged through shared memory by one pass of 256 threads");

__global__ void __launch_bounds__(256) bayes_weight_pack_batched_kernel(const PackEntry* __restrict__ table, const int n) {
    pdl_trigger();
    pdl_wait();
    __shared__ PackEntry e;
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].block0 <= (int)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    if (threadIdx.x < sizeof(PackEntry) / 4)
        reinterpret_cast<uint32_t*>(&e)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&table[lo])[threadIdx.x];
    __syncthreads();
    const int block = (int)blockIdx.x - e.block0;
    if (block < e.nblocks) weight_pack_block(e.p, e.NT, e.ntiles, e.nk, e.pack, e.fold_ln, e.vec, block);
}

__global__ void __launch_bounds__(128, 4) bayes_pointwise_tc_kernel(const BemBayesPointwiseParams p, const int NT, const int ntiles,
                                                                  const float* __restrict__ pack, const uint32_t tmem_cols) {
    pdl_trigger();
    pdl_wait();
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
                    const int64_t oi = (int64_t)co * p.P + pix;
                    float r = v[i] + b + (p.residual ? p.residual[(int64_t)img * p.cout * p.P + oi] : 0.f);
                    if (p.prelu_slope) r = r > 0.f ? r : r * p.prelu_slope[p.prelu_n > 1 ? co : 0];
                    out[oi] = r;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// persistent, warp-specialised form (the default): one CTA per SM walks (image, pixel tile, output-channel tile) items
//   warps 0-3   producers: raw x rows (cp.async, one 512-byte row per warp instruction) into a deep shared-memory ring
//   warp 4      MMA issuer: 3 x tcgen05.mma per K step, A operand from TMEM, B from shared memory
//   warp 5      TMA producer: packed weight tiles (ring) and the per-channel epilogue vectors
//   warps 6-13  transform: raw x tile -> (x - shift) split into tf32 hi / lo -> tcgen05.st into the A ring in TMEM
//               (pixel = TMEM lane, channel = column), two groups of four warps on alternate chunks; LayerNorm sums
//   warps 14-21 epilogue: tcgen05.ld -> rstd * (acc - mean * s_n) + t_n -> 128-byte coalesced stores
// so the loads of item i+1, the MMAs of item i and the stores of item i-1 overlap. Keeping A in TMEM takes its 16 KB of
// stores and 24 KB of MMA operand reads per chunk off shared memory, which otherwise bounds the kernel.
// TMEM columns: [0,192) and [192,384) accumulators, [384,512) A ring of 4 stages x (16 hi + 16 lo).
// Rows of A past the end of the image carry zero-filled or stale values: row m of D depends on row m of A only and
// those rows are never stored. Input channels past `cin` (last K chunk) are zeroed, they feed every output.
// ------------------------------------------------------------------------------------------------
// stage timeline (tools/trace_pointwise.py, env BEM_PW_TRACE=1): (tag, arg, SM clock) records of CTA 0, one region of
// TRACE_PER records per traced warp, plain stores (nothing on the critical path waits for them)
constexpr int TRACE_ROLES = 8, TRACE_PER = 2048;
__device__ uint4 g_trace[TRACE_ROLES * TRACE_PER];
struct Tracer {
    uint32_t n = 0;
    __device__ __forceinline__ void operator()(int on, int role, uint32_t tag, uint32_t arg) {
        if (on && blockIdx.x == 0 && n < TRACE_PER) g_trace[role * TRACE_PER + n++] = make_uint4(tag, arg, (uint32_t)clock64(), 1u);
    }
};

constexpr int P3_RS_MAX = 16;  // raw x stages  (TC_KC rows x 128 pixels x 4 B = 8 KB each), as many as fit
constexpr int P3_AS = 4;       // A stages in TMEM (16 hi + 16 lo columns each)
constexpr int P3_NMAX = 192;   // output channels per tile: two accumulators + the A ring fit the 512 TMEM columns
constexpr uint32_t P3_ACC_COLS = 192, P3_A_COL0 = 384;
constexpr int P3_BS_MAX = 8;   // B stages      (hi | lo, 2 * NT * TC_KC * 4 B each)
constexpr int P3_XW = 8;       // transform warps
constexpr int P3_EW = 8;       // epilogue warps (two per TMEM lane quarter, alternating 16-column groups)
constexpr int P3_PW = 4;       // activation producer warps (TC_KC / P3_PW rows of every chunk each)
constexpr int P3_W_MMA = P3_PW, P3_W_WGT = P3_PW + 1, P3_W_X0 = P3_PW + 2, P3_W_E0 = P3_W_X0 + P3_XW;   // first warp of each role
constexpr int P3_W_MMA2 = P3_W_E0 + P3_EW;   // second MMA issuer (DUAL)
constexpr int P3_THREADS = (P3_W_MMA2 + 1) * 32;
constexpr int P3_DUAL_NMAX = 96;             // DUAL: two partial accumulators of <= 96 columns per buffer

struct Ring {   // position in a ring of `n` stages and the phase bit of its mbarriers
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void next(uint32_t n) {
        if (++s == n) {
            s = 0;
            ph ^= 1;
        }
    }
};
constexpr uint32_t P3_RAW_BYTES = TC_KC * TC_M * 4;

struct P3Item {
    int tile, img, s_idx, npx;
    int64_t p0;
};
// item index -> (tile, pixel tile, image); 32-bit arithmetic (the host checks n_items < 2^31): 64-bit divisions cost
// ~1400 clk per item on the MMA warp's critical path
__device__ __forceinline__ P3Item p3_item(const BemBayesPointwiseParams& p, int64_t it64, int ntiles, int ptiles) {
    P3Item r;
    const uint32_t it = (uint32_t)it64;
    uint32_t q = it;
    r.tile = 0;
    if (ntiles > 1) {
        q = it / (uint32_t)ntiles;
        r.tile = (int)(it - q * (uint32_t)ntiles);
    }
    const uint32_t im = q / (uint32_t)ptiles;
    r.p0 = (int64_t)(q - im * (uint32_t)ptiles) * TC_M;
    r.img = (int)im;
    r.s_idx = p.n_samples > 1 ? (p.sample_interleave ? r.img % p.n_samples : r.img / (p.batch / p.n_samples)) : 0;
    r.npx = (int)min((int64_t)TC_M, p.P - r.p0);
    return r;
}

// EPI: what the epilogue adds to the per-channel affine — 0 nothing, 1 the skip connection (`residual`), 2 a PReLU. Separate
// instantiations: the store loop of the write-heavy layers is sensitive to every extra instruction and branch.
// DUAL (tiles of <= 96 output channels, >= 2 chunks): a `tcgen05.mma` costs its issuing thread ~50-100 clk whatever its N and
// a `tcgen05.commit` ~270 (tools/micro/mma_rate.cu: 105 clk per MMA at N = 16 .. 192 from one thread, the same per thread
// from two or four issuing warps), so with narrow tiles and many input channels one issuer bounds the pipeline at ~550 clk
// per 16-channel chunk. Two issuers take the chunks of even / odd running index (the transform groups' split) into their own
// accumulators (columns +0 / +96 of the buffer — no ordering between the two instruction streams is needed, and the sums
// stay deterministic); the epilogue adds the two partial sums. 160 -> 40: 56 -> 49 us, 320 -> 80: 45 -> 39 us.
template <bool LN, bool TRACE = false, int EPI = 0, bool DUAL = false>
__global__ void __launch_bounds__(P3_THREADS, 1) bayes_pointwise_tc3_kernel(const BemBayesPointwiseParams p, const int NT, const int ntiles,
                                                                          const int ptiles, const int64_t n_items,
                                                                          const float* __restrict__ pack, const float* __restrict__ vec,
                                                                          const uint32_t RS, const uint32_t BS, const int b_resident) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t b_bytes = (uint32_t)NT * TC_KC * 4;
    unsigned char* s_raw = smem;
    unsigned char* s_b = s_raw + RS * P3_RAW_BYTES;
    // B: a ring of BS stages, or (b_resident) every packed tile of the layer, BS = n_samples * ntiles * nk, loaded once
    float2* s_vec = reinterpret_cast<float2*>(s_b + BS * 2 * b_bytes);         // [n_samples * ntiles][NT] (s_n, t_n), resident
    float2* s_part = s_vec + p.n_samples * ntiles * NT;                        // [2 acc buffers][2 groups][128] (sum, sum sq)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 2 * TC_M);
    uint64_t* raw_full = bars;
    uint64_t* raw_empty = raw_full + P3_RS_MAX;
    uint64_t* a_full = raw_empty + P3_RS_MAX;
    uint64_t* a_empty = a_full + P3_AS;
    uint64_t* b_full = a_empty + P3_AS;
    uint64_t* b_empty = b_full + P3_BS_MAX;
    uint64_t* acc_full = b_empty + P3_BS_MAX;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* stats_full = acc_empty + 2;
    uint64_t* vec_full = stats_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(vec_full + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = (p.cin + TC_KC - 1) / TC_KC;
    Tracer tr;
    constexpr int trace_on = TRACE ? 1 : 0;   // the timeline build is its own instantiation: no trace predicates in the product kernel
    if (TRACE && tid == 0) tr(trace_on, 6, 40, 0);

    if (tid == 0) {
        for (int i = 0; i < (int)RS; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], P3_XW / 2); }
        for (int i = 0; i < P3_AS; ++i) { mbar_init(&a_full[i], P3_XW / 2); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < (b_resident ? 1 : (int)BS); ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(&vec_full[0], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], DUAL ? 2 : 1);
            mbar_init(&acc_empty[i], P3_EW);
            mbar_init(&stats_full[i], P3_XW);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    pdl_trigger();
    if (warp == P3_W_MMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (TRACE && tid == 0) tr(trace_on, 6, 41, 0);
    pdl_wait();   // barriers initialised and TMEM allocated while the previous kernel drains
    if (TRACE && tid == 0) tr(trace_on, 6, 42, 0);

    if (warp < P3_PW) {
        // ---------------- producers: activations (LDGSTS, 16 B per lane, one 512-byte row per instruction) ----------------
        // 512-byte bulk copies are bound by their per-copy overhead, and a cp.async-tracked mbarrier arrive serialises a
        // warp's copies behind it. So: plain cp.async groups and an ordinary arrive once a group has landed. Warp w owns
        // chunks g = w (mod P3_PW) — the loop is latency-bound (~600 clk per iteration), four of them interleave — and
        // keeps RS / P3_PW of them in flight. Rows past the last input channel and pixels past the image are zero-filled
        // through the src-size operand.
        const uint32_t lag = RS / P3_PW - 1;               // groups in flight per warp after which the oldest must land
        uint32_t g = 0, issued = 0;
        Ring r, done;                                       // ring positions of chunk g / of this warp's oldest pending chunk
        for (uint32_t i = 0; i < (uint32_t)warp; ++i) done.next(RS);
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            const P3Item w = p3_item(p, it, ntiles, ptiles);
            const int px = min(lane * 4, w.npx - 4);            // npx % 4 == 0; lanes past the end copy 0 bytes from a valid address
            const uint32_t pbytes = lane * 4 < w.npx ? 16u : 0u;
            const float* x = p.x + (int64_t)w.img * (p.x_img_stride ? p.x_img_stride : (int64_t)p.cin * p.P) + w.p0 + px;
            for (int kc = 0; kc < nk; ++kc, ++g, r.next(RS)) {
                if (g % P3_PW != (uint32_t)warp) continue;
                if (lane == 0) mbar_wait(&raw_empty[r.s], r.ph ^ 1, nullptr);
                __syncwarp();
                if (warp == 0 && lane == 0) tr(trace_on, 0, 2, g);
                const uint32_t dst = smem_u32(s_raw + r.s * P3_RAW_BYTES) + lane * 16;
                const int c0 = kc * TC_KC;
                {
                    if (c0 + TC_KC <= p.cin) {
                        const float* src = x + (int64_t)c0 * p.P;
#pragma unroll
                        for (int row = 0; row < TC_KC; ++row, src += p.P)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + row * (TC_M * 4)), "l"(src), "r"(pbytes) : "memory");
                    } else {
#pragma unroll
                        for (int row = 0; row < TC_KC; ++row) {
                            const float* src = x + (int64_t)min(c0 + row, p.cin - 1) * p.P;
                            const uint32_t sz = c0 + row < p.cin ? pbytes : 0u;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + row * (TC_M * 4)), "l"(src), "r"(sz) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (warp == 0 && lane == 0) tr(trace_on, 0, 1, g);
                if (++issued > lag) {
                    if (lag == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
                    else if (lag == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
                    else asm volatile("cp.async.wait_group 1;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&raw_full[done.s]);
                    if (warp == 0 && lane == 0) tr(trace_on, 0, 3, g);
                    for (int i = 0; i < P3_PW; ++i) done.next(RS);
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        for (uint32_t i = issued > lag ? issued - lag : 0; i < issued; ++i) {
            if (lane == 0) mbar_arrive(&raw_full[done.s]);
            for (int k = 0; k < P3_PW; ++k) done.next(RS);
        }
    } else if (warp == P3_W_WGT) {
        // ---------------- TMA producer: weights ----------------
        // The per-channel epilogue vectors are loaded once. So are the packed tiles when the whole layer fits
        // (b_resident: every level-0 layer) — consecutive items use the same weights; otherwise they stream through a ring.
        if (lane == 0) {
            const uint32_t vec_bytes = (uint32_t)(p.n_samples * ntiles * NT) * 8;
            mbar_arrive_expect_tx(&vec_full[0], vec_bytes);
            bulk_g2s(s_vec, vec, vec_bytes, &vec_full[0]);
            if (b_resident) {
                mbar_arrive_expect_tx(&b_full[0], BS * 2 * b_bytes);
                for (uint32_t i = 0; i < BS; ++i) bulk_g2s(s_b + i * 2 * b_bytes, pack + (int64_t)i * 2 * NT * TC_KC, 2 * b_bytes, &b_full[0]);
            } else {
                Ring r;
                for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                    const P3Item w = p3_item(p, it, ntiles, ptiles);
                    const float* bsrc = pack + ((int64_t)(w.s_idx * ntiles + w.tile) * nk) * 2 * NT * TC_KC;
                    for (int kc = 0; kc < nk; ++kc, r.next(BS)) {
                        mbar_wait(&b_empty[r.s], r.ph ^ 1, nullptr);
                        mbar_arrive_expect_tx(&b_full[r.s], 2 * b_bytes);
                        bulk_g2s(s_b + r.s * 2 * b_bytes, bsrc + (int64_t)kc * 2 * NT * TC_KC, 2 * b_bytes, &b_full[r.s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == P3_W_MMA2 && !DUAL) {
        // spare warp
    } else if (warp == P3_W_MMA || warp == P3_W_MMA2) {
        // ---------------- MMA issuer(s) ----------------
        const int issuer = warp == P3_W_MMA ? 0 : 1, istep = DUAL ? 2 : 1;
        // The whole warp walks the loop (uniform control flow keeps descriptors in uniform registers); one elected lane
        // issues. B descriptors are one base plus the stage / K-step offset in the 14-bit address field.
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        const uint64_t descB0 = make_desc(smem_u32(s_b), TC_LBO, TC_SBO);
        const uint32_t b_step = (2 * b_bytes) >> 4, b_lo = b_bytes >> 4;
        Ring ra, rb;
        uint32_t li = 0, g = 0;
        if (b_resident) mbar_wait(&b_full[0], 0, nullptr);
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
            const uint32_t buf = li & 1;
            if (b_resident) {   // stage index of this item's first tile within the resident set
                const P3Item w = p3_item(p, it, ntiles, ptiles);
                rb.s = (uint32_t)((w.s_idx * ntiles + w.tile) * nk);
            }
            mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1, nullptr);   // the epilogue has drained this accumulator
            if (lane == 0) tr(trace_on, 3 + 2 * issuer, 20 + 2 * issuer, li);
            const uint32_t d = tmem + buf * P3_ACC_COLS + (DUAL ? issuer * P3_DUAL_NMAX : 0);
            for (int kc = 0; kc < nk; ++kc, ++g, ra.next(P3_AS), rb.next(b_resident ? 0xffffffffu : BS)) {
                if (DUAL && (g & 1) != (uint32_t)issuer) continue;
                mbar_wait(&a_full[ra.s], ra.ph, nullptr);
                if (!b_resident) mbar_wait(&b_full[rb.s], rb.ph, nullptr);
                if (lane == 0) tr(trace_on, 3 + 2 * issuer, 21 + 2 * issuer, kc);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t aH = tmem + P3_A_COL0 + ra.s * (2 * TC_KC), aL = aH + TC_KC;
                    const uint64_t dBh = descB0 + (uint64_t)(rb.s * b_step), dBl = dBh + b_lo;
#pragma unroll
                    for (int ks = 0; ks < TC_KC / 8; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * ((2 * TC_LBO) >> 4));
                        umma_tf32_ts(d, aH + ks * 8, dBh + adv, idesc, ks ? 1u : (uint32_t)(kc >= istep));
                        umma_tf32_ts(d, aH + ks * 8, dBl + adv, idesc, 1);
                        umma_tf32_ts(d, aL + ks * 8, dBh + adv, idesc, 1);
                    }
                    umma_commit(&a_empty[ra.s]);
                    if (!b_resident) umma_commit(&b_empty[rb.s]);
                    if (kc + istep >= nk) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp < P3_W_E0) {
        // ---------------- transform: raw rows -> shifted, split -> A ring in TMEM ----------------
        // pixel = TMEM lane (a warp reaches lanes 32 * (warp % 4) ..). The two warps of a lane quarter take alternate
        // chunks (all 16 channels of a pixel each): the per-chunk handshakes of one group overlap the other group's.
        const int m = (warp & 3) * 32 + lane, grp = (warp - P3_W_X0) >> 2;
        const uint32_t a_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + P3_A_COL0;
        Ring rr, ra;
        uint32_t li = 0, g = 0;
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
            float s1 = 0.f, s2 = 0.f, shift = 0.f;
            if (LN) {   // per-pixel shift of the LayerNorm sums and of A: the pixel's first channel
                const P3Item w = p3_item(p, it, ntiles, ptiles);
                if (m < w.npx) shift = __ldg(p.x + (int64_t)w.img * (p.x_img_stride ? p.x_img_stride : (int64_t)p.cin * p.P) + w.p0 + m);
            }
            for (int kc = 0; kc < nk; ++kc, ++g, rr.next(RS), ra.next(P3_AS)) {
                if ((g & 1) != (uint32_t)grp) continue;
                mbar_wait(&raw_full[rr.s], rr.ph, nullptr);
                if ((warp & 3) == 0 && lane == 0) tr(trace_on, 1 + grp, 10 + grp, g);
                const float* raw = reinterpret_cast<const float*>(s_raw + rr.s * P3_RAW_BYTES) + m;
                float v[TC_KC];
#pragma unroll
                for (int e = 0; e < TC_KC; ++e) v[e] = raw[e * TC_M];
                const int left = p.cin - kc * TC_KC;                          // channels of this chunk that exist
                if (left < TC_KC) {
#pragma unroll
                    for (int e = 0; e < TC_KC; ++e) v[e] = e < left ? v[e] : shift;   // -> 0 after the shift
                }
                if (LN) {
#pragma unroll
                    for (int e = 0; e < TC_KC; ++e) {
                        v[e] -= shift;
                        s1 += v[e];
                        s2 = fmaf(v[e], v[e], s2);
                    }
                }
                float hi[TC_KC], lo[TC_KC];
#pragma unroll
                for (int e = 0; e < TC_KC; ++e) {
                    hi[e] = tf32_hi(v[e]);
                    lo[e] = v[e] - hi[e];
                }
                mbar_wait(&a_empty[ra.s], ra.ph ^ 1, nullptr);           // the MMAs that read this stage have completed
                if ((warp & 3) == 0 && lane == 0) tr(trace_on, 1 + grp, 12 + grp, g);
                tc_fence_after();
                tmem_st16(a_lane + ra.s * (2 * TC_KC), hi);
                tmem_st16(a_lane + ra.s * (2 * TC_KC) + TC_KC, lo);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[ra.s]);
                    mbar_arrive(&raw_empty[rr.s]);
                    if ((warp & 3) == 0) tr(trace_on, 1 + grp, 14 + grp, g);
                }
            }
            if (LN) {
                const uint32_t buf = li & 1;
                mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1, nullptr);   // the epilogue two items back has read its sums
                s_part[(buf * 2 + grp) * TC_M + m] = make_float2(s1, s2);
                __syncwarp();
                if (lane == 0) mbar_arrive(&stats_full[buf]);
            }
        }
    } else {
        // ---------------- epilogue ----------------
        const int ew = warp - P3_W_E0, q = warp & 3, half = ew >> 2, m = q * 32 + lane;   // TMEM lane quarter = warp % 4
        uint32_t li = 0;
        mbar_wait(&vec_full[0], 0, nullptr);
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
            const P3Item w = p3_item(p, it, ntiles, ptiles);
            const uint32_t buf = li & 1, par = (li >> 1) & 1;
            const int n0 = w.tile * NT, nvalid = min(NT, p.cout - n0);
            float rstd = 1.f, nmr = 0.f;
            if (LN) {
                mbar_wait(&stats_full[buf], par, nullptr);
                const float2 a = s_part[(buf * 2) * TC_M + m], b = s_part[(buf * 2 + 1) * TC_M + m];
                const float inv = 1.f / (float)p.cin;
                const float mean = (a.x + b.x) * inv;
                rstd = rsqrtf(fmaxf((a.y + b.y) * inv - mean * mean, 0.f) + p.ln_eps);
                nmr = -mean * rstd;
            }
            mbar_wait(&acc_full[buf], par, nullptr);
            if (ew == 0 && lane == 0) tr(trace_on, 4, 30, li);
            tc_fence_after();
            const bool valid = m < w.npx;
            const int64_t P = p.P;
            const int64_t obase = ((int64_t)w.img * p.cout + n0) * P + w.p0 + m;
            float* out = p.out + obase;
            const float* res = EPI == 1 ? p.residual + obase : nullptr;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * P3_ACC_COLS;
            const float2* sv = s_vec + (w.s_idx * ntiles + w.tile) * NT;
            const float* slope = p.prelu_slope;                 // PReLU after the conv (uniform branch), one slope or one per channel
            const int slope_step = p.prelu_n > 1 ? 1 : 0;
            const int ngrp = (nvalid + 15) >> 4;
            for (int gi = half; gi < ngrp; gi += 2) {
                const int c0 = gi * 16;
                float v[16];
                if constexpr (DUAL) {
                    float v2[16];
                    tmem_ld16x2(taddr + (uint32_t)c0, taddr + (uint32_t)(P3_DUAL_NMAX + c0), v, v2);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += v2[i];
                } else {
                    tmem_ld16(taddr + (uint32_t)c0, v);
                }
                float* o = out + (int64_t)c0 * P;
                float rv[EPI == 1 ? 16 : 1];
                if constexpr (EPI == 1) {   // skip connection: the loads go out together, ahead of the stores
#pragma unroll
                    for (int i = 0; i < 16; ++i) rv[i] = 0.f;
                    if (valid) {
                        const float* rp = res + (int64_t)c0 * P;
#pragma unroll
                        for (int i = 0; i < 16; ++i, rp += P)
                            if (c0 + i < nvalid) rv[i] = *rp;
                    }
                }
                // per-channel affine (+ skip connection) (+ PReLU: its own instantiation — even one uniform branch per group
                // here costs the write-heavy layers 4-6 %, a per-element predicate doubles their epilogue)
                if (c0 + 16 <= nvalid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 st = sv[c0 + i];
                        float r = LN ? fmaf(rstd, v[i], fmaf(nmr, st.x, st.y)) : v[i] + st.y;
                        if constexpr (EPI == 1) r += rv[i];
                        if constexpr (EPI == 2) r = r > 0.f ? r : r * slope[(n0 + c0 + i) * slope_step];
                        if (valid) *o = r;
                        o += P;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 st = sv[c0 + i];   // c0 + i < NT: groups of 16 within the NT-padded tile
                        float r = LN ? fmaf(rstd, v[i], fmaf(nmr, st.x, st.y)) : v[i] + st.y;
                        if constexpr (EPI == 1) r += rv[i];
                        if constexpr (EPI == 2) r = r > 0.f ? r : r * slope[min(n0 + c0 + i, p.cout - 1) * slope_step];
                        if (valid && c0 + i < nvalid) *o = r;
                        o += P;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (ew == 0 && lane == 0) tr(trace_on, 4, 31, li);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (TRACE && tid == 0) tr(trace_on, 6, 43, 0);
    if (warp == P3_W_MMA) tmem_dealloc(tmem, 512);
}

using P3Kernel = void (*)(const BemBayesPointwiseParams, int, int, int, int64_t, const float*, const float*, uint32_t, uint32_t, int);
template <bool LN, bool DUAL>
static P3Kernel p3_pick(int variant) {   // 0 plain, 1 skip connection, 2 PReLU, 3 plain + timeline
    switch (variant) {
        case 1: return bayes_pointwise_tc3_kernel<LN, false, 1, DUAL>;
        case 2: return bayes_pointwise_tc3_kernel<LN, false, 2, DUAL>;
        case 3: return bayes_pointwise_tc3_kernel<LN, true, 0, DUAL>;
        default: return bayes_pointwise_tc3_kernel<LN, false, 0, DUAL>;
    }
}

static void tc_tiling(int cin, int cout, int nmax, int& ntiles, int& NT, int& nk) {
    // output-channel tiling: as few tiles as possible, each a multiple of 16 and at most `nmax` channels
    ntiles = (cout + nmax - 1) / nmax;
    NT = ((cout + ntiles - 1) / ntiles + 15) / 16 * 16;
    if (NT < 16) NT = 16;
    nk = (cin + TC_KC - 1) / TC_KC;
}

// workspace = [packed tiles | per-(sample, tile) epilogue vectors]
static int64_t tc_pack_floats(int n_samples, int cin, int cout, int nmax) {
    int ntiles, NT, nk;
    tc_tiling(cin, cout, nmax, ntiles, NT, nk);
    return (int64_t)n_samples * ntiles * nk * 2 * NT * TC_KC;
}
static int64_t tc_workspace_floats(int n_samples, int cin, int cout, int nmax) {
    int ntiles, NT, nk;
    tc_tiling(cin, cout, nmax, ntiles, NT, nk);
    return tc_pack_floats(n_samples, cin, cout, nmax) + (int64_t)n_samples * ntiles * 2 * NT;
}
int64_t bayes_pointwise_tc_workspace(int n_samples, int cin, int cout) {
    // enough for either tiling (persistent kernel: tiles of <= 192 channels; unaligned-input kernel: <= 256)
    return std::max(tc_workspace_floats(n_samples, cin, cout, P3_NMAX), tc_workspace_floats(n_samples, cin, cout, TC_NMAX)) *
           (int64_t)sizeof(float);
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// which kernel a call takes, its tiling, and where its packed tiles / epilogue vectors live in the workspace
struct TcPlan {
    int ntiles, NT, nk, pack_blocks, vec_blocks;
    int64_t ptiles;
    bool persistent;
    float* pack;
    float* vec;
};
static int tc_plan(const BemBayesPointwiseParams& p, TcPlan& t) {
    const int64_t need = bayes_pointwise_tc_workspace(p.n_samples, p.cin, p.cout);
    if (!p.workspace || p.workspace_bytes < need || (reinterpret_cast<uintptr_t>(p.workspace) & 15)) return BEM_ERR_WORKSPACE;
    t.ptiles = (p.P + TC_M - 1) / TC_M;
    if (t.ptiles > 65535 || p.batch > 65535) return BEM_ERR_UNSUPPORTED;
    t.pack = reinterpret_cast<float*>(p.workspace);
    // the persistent kernel moves x with 16-byte cp.async: rows must start and end on 16-byte boundaries
    static const int force_v2 = env_int("BEM_PW_V2", 0);
    const bool aligned = (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && p.P % 4 == 0 && p.x_img_stride % 4 == 0;
    tc_tiling(p.cin, p.cout, P3_NMAX, t.ntiles, t.NT, t.nk);
    // the persistent kernel keeps the epilogue vectors of every (sample, tile) in shared memory
    // (S-batched Monte-Carlo forwards: S weight sets x up to 1280 output channels = 40 KB at S = 4; the raw-stage count shrinks to fit)
    t.persistent = aligned && !force_v2 && (int64_t)p.n_samples * t.ntiles * t.NT * 8 <= 64 * 1024;
    if (!t.persistent) tc_tiling(p.cin, p.cout, TC_NMAX, t.ntiles, t.NT, t.nk);
    t.vec = t.pack + tc_pack_floats(p.n_samples, p.cin, p.cout, t.persistent ? P3_NMAX : TC_NMAX);
    t.pack_blocks = p.n_samples * t.ntiles * t.nk;
    t.vec_blocks = p.n_samples * ((p.cout + 7) / 8);
    return BEM_OK;
}

int64_t bayes_pointwise_pack_table_bytes(int n) { return (int64_t)n * (int64_t)sizeof(PackEntry); }

int bayes_pointwise_pack_table(const BemBayesPointwiseParams* params, int n, void* table_host, int32_t* total_blocks) {
    PackEntry* tab = reinterpret_cast<PackEntry*>(table_host);
    int64_t blocks = 0;
    for (int i = 0; i < n; ++i) {
        TcPlan t;
        const int rc = tc_plan(params[i], t);
        if (rc != BEM_OK) return rc;
        PackEntry& e = tab[i];
        e.p = params[i];
        e.pack = t.pack;
        e.vec = t.vec;
        e.NT = t.NT;
        e.ntiles = t.ntiles;
        e.nk = t.nk;
        e.fold_ln = t.persistent ? 1 : 0;
        e.block0 = (int32_t)blocks;
        e.nblocks = t.pack_blocks + (t.persistent ? t.vec_blocks : 0);
        blocks += e.nblocks;
        if (blocks > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    }
    *total_blocks = (int32_t)blocks;
    return BEM_OK;
}

int bayes_pointwise_pack_run(const void* table_dev, int n, int total_blocks, cudaStream_t stream) {
    launch_pdl(bayes_weight_pack_batched_kernel, dim3(total_blocks), dim3(256), 0, stream, reinterpret_cast<const PackEntry*>(table_dev), n);
    return (int)cudaGetLastError();
}

int bayes_pointwise_tc_launch(const BemBayesPointwiseParams& p, cudaStream_t stream) {
    TcPlan t;
    const int rc = tc_plan(p, t);
    if (rc != BEM_OK) return rc;
    const int64_t ptiles = t.ptiles;
    const int ntiles = t.ntiles, NT = t.NT, nk = t.nk, pack_blocks = t.pack_blocks, vec_blocks = t.vec_blocks;
    const bool persistent = t.persistent;
    float* pack = t.pack;
    float* vec = t.vec;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    static const int trace_on = env_int("BEM_PW_TRACE", 0);   // record CTA 0's stage timeline (tools/trace_pointwise.py)
    if (persistent) {
        // shared-memory plan: the layer's packed tiles resident when they fit in 120 KB, else a B ring of >= 3 stages
        // (up to ~48 KB); the epilogue vectors resident; the rest goes to raw x stages in flight
        const int b_stage = 2 * NT * TC_KC * 4;
        const int all_stages = p.n_samples * ntiles * nk;
        const int vec_bytes = p.n_samples * ntiles * NT * 8;
        const int b_resident = (int64_t)all_stages * b_stage <= 120 * 1024;
        const int BS = b_resident ? all_stages : std::max(3, std::min(P3_BS_MAX, (48 * 1024) / b_stage));
        const int fixed = BS * b_stage + vec_bytes + 2 * 2 * TC_M * 8 + (2 * P3_RS_MAX + 2 * P3_AS + 2 * P3_BS_MAX + 8) * 8 + 16;
        const int RS = std::max(2 * P3_PW, std::min(P3_RS_MAX, (int)((220 * 1024 - fixed) / (int)P3_RAW_BYTES)) / P3_PW * P3_PW);
        const int smem_bytes = RS * (int)P3_RAW_BYTES + fixed;
        if (smem_bytes > 227 * 1024) return BEM_ERR_UNSUPPORTED;
        static int attr3[64][2][2][4] = {}, sms[64] = {0};
        if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (!p.prepacked) launch_pdl(bayes_weight_pack_kernel, dim3(pack_blocks + vec_blocks), dim3(256), 0, stream, p, NT, ntiles, nk, pack, 1, vec);
        const int64_t n_items = (int64_t)p.batch * ptiles * ntiles;
        if (n_items >= (1ll << 31)) return BEM_ERR_UNSUPPORTED;
        const int grid = (int)std::min<int64_t>(n_items, sms[dev]);
        if (p.residual && p.prelu_slope) return BEM_ERR_UNSUPPORTED;   // no caller combines them (the PReLU layers have no skip)
        // two MMA issuers pay off when the issue rate bounds the tile: many input channels into a narrow tile (A/B knob: BEM_PW_NO_DUAL)
        static const int no_dual = env_int("BEM_PW_NO_DUAL", 0);
        const int ln = p.ln_gamma ? 1 : 0, dual = (NT <= P3_DUAL_NMAX && nk >= 4 && !no_dual) ? 1 : 0;
        const int variant = p.residual ? 1 : p.prelu_slope ? 2 : trace_on ? 3 : 0;   // the timeline build has the plain epilogue only
        const P3Kernel kernel = ln ? (dual ? p3_pick<true, true>(variant) : p3_pick<true, false>(variant))
                                   : (dual ? p3_pick<false, true>(variant) : p3_pick<false, false>(variant));
        if (attr3[dev][ln][dual][variant] < smem_bytes) {
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (e != cudaSuccess) return (int)e;
            attr3[dev][ln][dual][variant] = smem_bytes;
        }
        launch_pdl(kernel, dim3(grid), dim3(P3_THREADS), smem_bytes, stream, p, NT, ntiles, (int)ptiles, n_items, pack, vec, RS, BS, b_resident);
        return (int)cudaGetLastError();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NT) tmem_cols <<= 1;
    const int smem_bytes = 2 * (2 * TC_M * TC_KC * 4 + 2 * NT * TC_KC * 4) + 48 + 2 * p.cin * 4 + 64;
    if (smem_bytes > 227 * 1024) return BEM_ERR_UNSUPPORTED;
    static int attr_set[64] = {0};
    if (attr_set[dev] < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(bayes_pointwise_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = smem_bytes;
    }
    if (!p.prepacked) launch_pdl(bayes_weight_pack_kernel, dim3(pack_blocks), dim3(256), 0, stream, p, NT, ntiles, nk, pack, 0, vec);
    dim3 grid((unsigned)ntiles, (unsigned)ptiles, (unsigned)p.batch);
    launch_pdl(bayes_pointwise_tc_kernel, dim3(grid), dim3(128), smem_bytes, stream, p, NT, ntiles, pack, tmem_cols);
    return (int)cudaGetLastError();
}

}  // namespace bem

// not part of the ABI: reads (and clears) the debug timeline of the persistent pointwise kernel (tools/trace_pointwise.py)
extern "C" int bem_dbg_pointwise_trace(unsigned int* out, int max_records) {
    const int total = bem::TRACE_ROLES * bem::TRACE_PER;
    if (max_records < total) return -1;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, bem::g_trace, (size_t)total * sizeof(uint4));
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, bem::g_trace);
    cudaMemset(sym, 0, (size_t)total * sizeof(uint4));
    return total;
}
