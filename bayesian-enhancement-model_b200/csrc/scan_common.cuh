// scan_common.cuh — device helpers shared by the sm_100a selective-scan kernels.
//
// Building blocks (all hand-written, no CUB):
//   * mbarrier + cp.async.bulk (TMA 1-D bulk copies) wrappers: global -> shared staging with transaction
//     barriers, shared -> global bulk stores with bulk_group completion
//   * the scan monoid of the SSM recurrence h_t = a_t h_{t-1} + b_t  (reference: SSMScanOp,
//     kernels/selective_scan/csrc/selective_scan/selective_scan_common.h:91-96) as warp-shuffle scans
//   * a decoupled look-back over per-(row, chunk, state) descriptors so the sequence dimension L is split
//     across CTAs (the reference walks its 2048-chunks serially inside one CTA,
//     cusoflex/selective_scan_fwd_kernel_oflex.cuh:110)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bem {

constexpr unsigned FULL = 0xffffffffu;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (ns): the warp sleeps in hardware instead of re-issuing the probe
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box. On timeout the error word is set and the wait
// returns; the results are then garbage and the host reports BEM failure from the error word.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* err) {
    if (mbar_try_wait(bar, parity)) return;
    int tries = 0;
    while (!mbar_try_wait(bar, parity)) {   // each probe suspends the warp in hardware for a bounded time
        if (++tries > (1 << 24)) {
            if (err) atomicOr(err, 1u);
            return;
        }
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Decay of the recurrence, returned as e = exp(y) - 1 with y = delta * A <= 0, and applied as h <- fma(e, h, h) + b.
// Why not a = exp(y): the recurrence multiplies ~1/(1-a) consecutive decays. (i) A correlated 2-ulp error of
// ex2.approx on near-unit decays is amplified to ~1e-4 of the state in slow channels (|A| small); (ii) even a
// correctly rounded a = 1 - 4e-5 carries only ~10 significant bits of the decay RATE. For y > -1/8, e comes from a
// degree-5 Taylor polynomial of expm1 (truncation < 4.5e-8 relative), so the rate keeps full fp32 precision; faster
// decays have a short memory and use the single MUFU instruction (a - 1 is exact for a in [0.5, 1]).
// kAccurate = false (16-bit inputs, 1e-2 tolerance) always takes the MUFU branch.
template <bool kAccurate>
__device__ __forceinline__ float decay_m1(float y) {
    const float fast = ex2_approx(y * kLog2e) - 1.f;
    if constexpr (!kAccurate) return fast;
    float p = fmaf(y, 1.f / 120.f, 1.f / 24.f);
    p = fmaf(y, p, 1.f / 6.f);
    p = fmaf(y, p, 0.5f);
    const float e = fmaf(y * y, p, y);
    return y > -0.125f ? e : fast;
}
// one step of the affine map composition with the decay given as e = a - 1: (P, V) <- (a P, a V + b)
__device__ __forceinline__ void decay_step(float e, float b, float& P, float& V) {
    V = fmaf(e, V, V) + b;
    P = fmaf(e, P, P);
}

// softplus(x) = max(x,0) + log1p(exp(-|x|)); matches the reference's `x <= 20 ? log1pf(expf(x)) : x`
// (cusoflex/selective_scan_fwd_kernel_oflex.cuh:124-126, F.softplus threshold 20): beyond 20 the correction term is below
// 2.1e-9. log1p(z) must keep its RELATIVE accuracy for small z: x = delta + bias is very negative for channels whose dt
// sits near dt_init_floor (z ~ 1e-4), and forming 1.f + z rounds z to 6e-8 absolute = 6e-4 of such a delta, which the
// backward pass multiplies into ddelta and dA. Below z = 2^-4 the alternating series is used instead; at and above it the
// rounding of 1 + z costs at most 2^-24 / 2^-4 = 1e-6 relative.
__device__ __forceinline__ float log1p_small(float z) {
    // z in [0, 2^-4): z * (1 - z/2 + z^2/3 - z^3/4 + z^4/5 - z^5/6), truncation < z^6/7 < 9e-9 relative
    float p = fmaf(z, -1.f / 6.f, 0.2f);
    p = fmaf(z, p, -0.25f);
    p = fmaf(z, p, 1.f / 3.f);
    p = fmaf(z, p, -0.5f);
    p = fmaf(z, p, 1.f);
    return z * p;
}
__device__ __forceinline__ float softplus_f(float x) {
    const float z = ex2_approx(-fabsf(x) * kLog2e);
    const float big = kLn2 * lg2_approx(1.f + z);          // z >= 2^-4: rounding of 1 + z costs <= 1e-6 relative
    return fmaxf(x, 0.f) + (z < 0.0625f ? log1p_small(z) : big);
}
// ------------------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 execute two fp32 operations per issued instruction). The scan kernels are
// bound by instruction issue, and two thirds of their instructions are fp32 arithmetic on independent elements, so the element
// math is evaluated for TWO positions at a time. Every packed operation is the same IEEE operation as its scalar twin
// (fma.rn / mul.rn / add.rn per half), and the helpers below mirror softplus_f / decay_m1 operation for operation: results
// are bit-identical to the scalar forms.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 r, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
__device__ __forceinline__ f32x2 splat2(float v) { return pk2(v, v); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// softplus_f of two values (same operations per half)
__device__ __forceinline__ f32x2 softplus2(float x0, float x1) {
    const float z0 = ex2_approx(-fabsf(x0) * kLog2e), z1 = ex2_approx(-fabsf(x1) * kLog2e);
    const f32x2 z = pk2(z0, z1);
    float u0, u1;
    upk2(add2(splat2(1.f), z), u0, u1);
    const f32x2 big = mul2(splat2(kLn2), pk2(lg2_approx(u0), lg2_approx(u1)));
    f32x2 p = fma2(z, splat2(-1.f / 6.f), splat2(0.2f));
    p = fma2(z, p, splat2(-0.25f));
    p = fma2(z, p, splat2(1.f / 3.f));
    p = fma2(z, p, splat2(-0.5f));
    p = fma2(z, p, splat2(1.f));
    const f32x2 sm = mul2(z, p);
    float s0, s1, b0, b1;
    upk2(sm, s0, s1);
    upk2(big, b0, b1);
    return add2(pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)), pk2(z0 < 0.0625f ? s0 : b0, z1 < 0.0625f ? s1 : b1));
}
// decay_m1<true> of two values
__device__ __forceinline__ f32x2 decay_m1_2(f32x2 y) {
    float y0, y1, t0, t1;
    upk2(y, y0, y1);
    upk2(mul2(y, splat2(kLog2e)), t0, t1);
    const f32x2 fast = add2(pk2(ex2_approx(t0), ex2_approx(t1)), splat2(-1.f));
    f32x2 p = fma2(y, splat2(1.f / 120.f), splat2(1.f / 24.f));
    p = fma2(y, p, splat2(1.f / 6.f));
    p = fma2(y, p, splat2(0.5f));
    const f32x2 e = fma2(mul2(y, y), p, y);
    float e0, e1, f0, f1;
    upk2(e, e0, e1);
    upk2(fast, f0, f1);
    return pk2(y0 > -0.125f ? e0 : f0, y1 > -0.125f ? e1 : f1);
}

// d softplus / dx = sigmoid(x); the reference switches to 1 above the threshold
// (cusoflex/selective_scan_bwd_kernel_oflex.cuh:250-255)
__device__ __forceinline__ float softplus_grad_f(float x) {
    return x <= 20.f ? __fdividef(1.f, 1.f + ex2_approx(-x * kLog2e)) : 1.f;
}

// 128-bit descriptor access that is a single transaction at L2 (bypasses L1)
__device__ __forceinline__ uint4 ld_desc(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint4* p, float P, float V, uint32_t status) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(P)), "r"(__float_as_uint(V)),
                 "r"(status), "r"(0u)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// element type helpers
// ------------------------------------------------------------------------------------------------
template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
    static constexpr int kPerVec = 4;
    __device__ static __forceinline__ float to_f(float v) { return v; }
    __device__ static __forceinline__ float from_f(float v) { return v; }
};
template <> struct ElemTraits<__half> {
    static constexpr int kPerVec = 8;
    __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct ElemTraits<__nv_bfloat16> {
    static constexpr int kPerVec = 8;
    __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// ITEMS consecutive elements from 16-byte aligned shared memory -> fp32 registers (LDS.128)
template <typename T, int ITEMS>
__device__ __forceinline__ void lds_items(const T* __restrict__ src, float (&dst)[ITEMS]) {
    constexpr int V = ElemTraits<T>::kPerVec;
    static_assert(ITEMS % V == 0, "ITEMS must be a multiple of the 128-bit vector width");
#pragma unroll
    for (int v = 0; v < ITEMS / V; ++v) {
        const uint4 raw = reinterpret_cast<const uint4*>(src)[v];
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k) dst[v * V + k] = ElemTraits<T>::to_f(e[k]);
    }
}
// fp32 registers -> ITEMS consecutive elements in 16-byte aligned shared memory (STS.128)
template <typename T, int ITEMS>
__device__ __forceinline__ void sts_items(T* __restrict__ dst, const float (&src)[ITEMS]) {
    constexpr int V = ElemTraits<T>::kPerVec;
    static_assert(ITEMS % V == 0, "ITEMS must be a multiple of the 128-bit vector width");
#pragma unroll
    for (int v = 0; v < ITEMS / V; ++v) {
        uint4 raw;
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k) e[k] = ElemTraits<T>::from_f(src[v * V + k]);
        reinterpret_cast<uint4*>(dst)[v] = raw;
    }
}

// ------------------------------------------------------------------------------------------------
// scan monoid: an affine map h -> P*h + V. combine(first, second) applies `first` then `second`.
// ------------------------------------------------------------------------------------------------
// inclusive scan over the warp in lane order (lane 0 is applied first)
__device__ __forceinline__ void warp_scan_fwd(float& P, float& V, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Pp = __shfl_up_sync(FULL, P, o);
        const float Vp = __shfl_up_sync(FULL, V, o);
        if (lane >= o) {
            V = fmaf(P, Vp, V);
            P = P * Pp;
        }
    }
}
// inclusive scan over the warp in REVERSE lane order (lane 31 is applied first)
__device__ __forceinline__ void warp_scan_rev(float& P, float& V, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Pp = __shfl_down_sync(FULL, P, o);
        const float Vp = __shfl_down_sync(FULL, V, o);
        if (lane + o < 32) {
            V = fmaf(P, Vp, V);
            P = P * Pp;
        }
    }
}

// Descriptor status. The status word carries the launch epoch (workspace header word 2, incremented by the last ticket
// drawer of every launch): a descriptor is valid only if it was written by THIS launch, so the workspace never has to
// be cleared between launches (a ~1.6 MB memset node per scan otherwise) — it is zero-filled once by its owner.
enum : uint32_t { DESC_EMPTY = 0, DESC_READY = 1 };
__device__ __forceinline__ uint32_t desc_tag(uint32_t epoch, uint32_t kind) { return (epoch << 2) | kind; }
__device__ __forceinline__ uint32_t desc_kind(uint32_t word, uint32_t epoch) { return (word >> 2) == (epoch & 0x3fffffffu) ? (word & 3u) : DESC_EMPTY; }
constexpr int kAnchor = 16;   // every kAnchor-th tile of a row also publishes its INCLUSIVE composition

// Deterministic decoupled look-back.
// Tile c of a row needs the composition of all tiles the flowing value passes through BEFORE it: step = -1 -> tiles
// c-1 ... 0 (forward scan), step = +1 -> tiles c+1 ... nt-1 (reverse scan). Every tile publishes its own AGGREGATE
// (descriptor array `agg`) as soon as its local scan is done, before it looks back itself; tiles whose distance from
// the sequence start (end, for the reverse scan) is a multiple of kAnchor additionally publish their INCLUSIVE
// composition (array `incl`) once they know it. Tile c combines, in a FIXED order, the aggregates of the tiles back
// to an anchor that lies kAnchor .. 2*kAnchor-1 tiles behind it with that anchor's inclusive value: at most 31
// descriptors, one per lane, one round trip. Choosing an anchor at least kAnchor tiles back means that in steady state
// it belongs to an earlier wave of tiles and is long finished, so nothing ever waits on a concurrently running tile's
// look-back (decoupled); and because the association tree depends only on c, results are bit-reproducible (the
// classic look-back stops at whichever predecessor happens to be inclusive already and is not).
// dist = number of tiles before c in flow order: c (forward) or nt-1-c (reverse).
struct LookbackPlan {
    int nlanes;         // descriptors to read (0 = nothing before this tile)
    bool anchor_incl;   // the last participating lane reads an anchor's inclusive descriptor (else tile 0's aggregate)
    bool publish_agg;   // a later tile will read our aggregate
    bool publish_incl;  // we are an anchor that later tiles will read
};
__device__ __forceinline__ LookbackPlan lookback_plan(int dist, int ntiles) {
    LookbackPlan pl;
    const int anchor = dist >= kAnchor ? (dist / kAnchor - 1) * kAnchor : -1;   // in [dist-2*kAnchor+1, dist-kAnchor]
    pl.anchor_incl = anchor >= 0;
    pl.nlanes = anchor >= 0 ? dist - anchor : dist;   // tiles dist-1 ... anchor (or ... 0)
    pl.publish_agg = dist + 1 < ntiles;
    pl.publish_incl = (dist % kAnchor) == 0 && dist + kAnchor < ntiles;
    return pl;
}
// lane j inspects the tile j+1 steps before c in flow order; with an anchor, the last participating lane reads the
// anchor's inclusive descriptor. Returns this lane's descriptor pointer (nullptr for idle lanes).
__device__ __forceinline__ const uint4* lookback_addr(const uint4* agg0, const uint4* incl0, int64_t dstride, int c, int step,
                                                      const LookbackPlan& pl, int lane) {
    if (lane >= pl.nlanes) return nullptr;
    const int idx = c + step * (lane + 1);
    return ((pl.anchor_incl && lane == pl.nlanes - 1) ? incl0 : agg0) + (int64_t)idx * dstride;
}
__device__ __forceinline__ uint4 lookback_prefetch(const uint4* addr) {
    return addr ? ld_desc(addr) : make_uint4(0u, 0u, DESC_READY, 0u);
}
// `first` is the result of an early lookback_prefetch of this lane's descriptor (issued before the tile's arithmetic).
__device__ __forceinline__ float2 lookback_finish(const uint4* addr, uint4 first, int nlanes, int lane, unsigned int* err, uint32_t epoch) {
    uint4 v = first;
    // warp-convergent poll: one instruction stream for the whole warp, only the lanes still waiting reload
    // (per-lane spin loops would diverge into up to 31 independent loops that hog the scheduler's issue slots)
    bool pending = addr != nullptr && desc_kind(v.z, epoch) == DESC_EMPTY;
    int spins = 0;
    while (__any_sync(FULL, pending)) {
        if (pending) {
            v = ld_desc(addr);
            pending = desc_kind(v.z, epoch) == DESC_EMPTY;
        }
        if (++spins > 16) __nanosleep(64);
        if (spins > (1 << 22)) {   // watchdog, see mbar_wait
            if (lane == 0) atomicOr(err, 2u);
            break;
        }
    }
    __syncwarp();
    float P = 1.f, V = 0.f;
    if (lane < nlanes) {
        P = __uint_as_float(v.x);
        V = __uint_as_float(v.y);
    }
    // fold lanes: lane i ends with tiles [i, 32) where higher lanes (further away) are applied first
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Pp = __shfl_down_sync(FULL, P, o);
        const float Vp = __shfl_down_sync(FULL, V, o);
        if (lane + o < 32) {
            V = fmaf(P, Vp, V);
            P = P * Pp;
        }
    }
    return make_float2(__shfl_sync(FULL, P, 0), __shfl_sync(FULL, V, 0));
}

// Classic (timing-dependent) decoupled look-back, kept for A/B measurements: walks back over windows of 32 tiles and
// stops at the first tile that has published an inclusive value. `agg0` holds status 1 = aggregate, 2 = inclusive.
__device__ __forceinline__ float2 lookback_dynamic(const uint4* desc0, int64_t dstride, int c, int nchunks, int step, int lane,
                                                   unsigned int* err, uint32_t epoch) {
    float runP = 1.f, runV = 0.f;
    int j = c + step;
    while (true) {
        const int idx = j + step * lane;
        const bool inside = idx >= 0 && idx < nchunks;
        float P = 1.f, V = 0.f;
        uint32_t st = 2u;
        if (inside) st = 0u;
        unsigned incl, need;
        int spins = 0;
        while (true) {
            if (inside && st == 0u) {
                const uint4 v = ld_desc(desc0 + (int64_t)idx * dstride);
                st = desc_kind(v.z, epoch);
                P = __uint_as_float(v.x);
                V = __uint_as_float(v.y);
            }
            incl = __ballot_sync(FULL, st == 2u);
            const unsigned ready = __ballot_sync(FULL, st != 0u);
            need = incl ? ((2u << (__ffs(incl) - 1)) - 1u) : FULL;
            if ((ready & need) == need) break;
            if (++spins > 16) __nanosleep(64);
            if (spins > (1 << 22)) {
                if (lane == 0) atomicOr(err, 2u);
                break;
            }
        }
        if (!((need >> lane) & 1u)) {
            P = 1.f;
            V = 0.f;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float Pp = __shfl_down_sync(FULL, P, o);
            const float Vp = __shfl_down_sync(FULL, V, o);
            if (lane + o < 32) {
                V = fmaf(P, Vp, V);
                P = P * Pp;
            }
        }
        const float Pw = __shfl_sync(FULL, P, 0), Vw = __shfl_sync(FULL, V, 0);
        runV = fmaf(runP, Vw, runV);
        runP = runP * Pw;
        if (incl) break;
        j += step * 32;
    }
    return make_float2(runP, runV);
}

// ------------------------------------------------------------------------------------------------
// work decomposition shared by fwd/bwd
// ------------------------------------------------------------------------------------------------
struct TileCoord {
    int c;       // tile index along L
    int b, g;    // batch, B/C group
    int row0;    // first channel row (within the group) of this step
    int nrows;   // rows in this step (<= NW); < 0 marks the end of work
    int len;     // valid positions in this tile
    int aux0, aux1;
    uint32_t epoch;   // launch epoch of the look-back descriptors (see desc_tag)
};

}  // namespace bem
