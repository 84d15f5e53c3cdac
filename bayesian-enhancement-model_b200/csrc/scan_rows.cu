// scan_rows.cu — selective scan for dstate >= 2: one CTA per channel row, the sequence walked chunk by chunk.
//
// Replaces, for the multi-state configurations (VMamba's d_state 16; the reference builds up to MAX_DSTATE 256,
// selective_scan_oflex.cpp:190), selective_scan_fwd_kernel / selective_scan_bwd_kernel
// (kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_fwd_kernel_oflex.cuh:73-236,
// selective_scan_bwd_kernel_oflex.cuh:73-289). Same recurrence and gradients as scan_fwd.cu / scan_bwd.cu.
//
// Why a second organisation. With N states per row the work per position is N times that of the dstate-1 kernels while the
// bytes barely change (u, delta, out per row; B, C are shared by the Dg rows of a group and stay in L2): these shapes are
// bound by instruction issue and by the latency of the per-state scan chain, not by HBM. The look-back kernels spend one
// descriptor exchange per (row, tile, STATE); here a row's states live in ONE CTA that walks the chunks in order, so the
// state crossing a chunk boundary is a register-to-shared-memory hand-off and there is no inter-CTA protocol at all.
//
//   * 8 warps per CTA; warp w owns states w, w + 8, w + 16, ... of the row (any dstate <= 256 runs, two states per warp at
//     d_state 16). A lane holds ITEMS consecutive positions of the chunk (chunk = the carry chunk of `x`, bem_scan_chunk_len).
//   * What every state shares is computed once per chunk by the whole CTA ("pre-pass", one or two positions per thread):
//     softplus(delta + bias), delta * u, D * u; the global loads of the next chunk are issued a chunk ahead.
//   * B_n / C_n slices are private to the owning warp: staged with cp.async into two per-warp slots, two work items ahead.
//   * Per (chunk, state): local scan over the lane's positions, warp-shuffle scan, carry from the previous chunk, C_n . h.
//     The states' contributions are summed per warp in registers, across warps through shared memory (one barrier pair per chunk).
//   * Backward: chunks in reverse order, the forward state at the chunk start comes from `x` (written by the forward pass, as
//     in the reference, bwd_kernel_oflex.cuh:200), the adjoint crossing the chunk boundary is carried like the forward state.
//     dB / dC (summed over the Dg rows of a group) leave as 128-bit vector reductions (REDG.F32x4), a quarter of the
//     reference's 2 N L scalar atomics per row (bwd_kernel_oflex.cuh:224-237).
#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

namespace {

constexpr int kRowsWarps = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One warp stages `len` valid elements of a row slice into CL shared-memory slots; the tail [len, CL) is zero-filled so that
// positions beyond the sequence end are identities of the recurrence (b = 0) whatever the slot held before.
template <typename T>
__device__ __forceinline__ void stage_slice(T* dst, const T* src, int len, int CL, int lane) {
    constexpr int PER = 16 / (int)sizeof(T);
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nfull = len / PER;
        for (int v = lane; v < nfull; v += 32) cp_async16(dst + v * PER, src + v * PER);
        done = nfull * PER;
    }
    for (int e = done + lane; e < CL; e += 32) dst[e] = e < len ? src[e] : ElemTraits<T>::from_f(0.f);
}

// Fast path of the above for a full chunk of a 16-byte aligned row: ITEMS * sizeof(T) / 16 vectors per lane, no loops.
template <typename T, int ITEMS>
__device__ __forceinline__ void stage_full(T* dst, const T* src, int lane) {
    constexpr int PER = 16 / (int)sizeof(T);
    constexpr int NV = ITEMS / PER;
#pragma unroll
    for (int q = 0; q < NV; ++q) cp_async16(dst + (lane + 32 * q) * PER, src + (lane + 32 * q) * PER);
}
// Loads whose position in the instruction stream matters (issued a chunk ahead of their use): as volatile asm the compiler
// cannot sink them down to the use to save registers.
__device__ __forceinline__ float ldg_early(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_early(const __half* p) {
    unsigned short v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return __half2float(__ushort_as_half(v));
}
__device__ __forceinline__ float ldg_early(const __nv_bfloat16* p) {
    unsigned short v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return __bfloat162float(__ushort_as_bfloat16(v));
}
__device__ __forceinline__ bool aligned16(const void* p, int64_t stride_bytes) {
    return ((reinterpret_cast<uintptr_t>(p) | (uintptr_t)stride_bytes) & 15) == 0;
}

}  // namespace

// =====================================================================================================================
// forward
// =====================================================================================================================
template <typename T, typename OutT, int ITEMS>
__global__ void __launch_bounds__(kRowsWarps * 32, 3) scan_rows_fwd_kernel(const ScanFwdArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NW = kRowsWarps, NT = NW * 32, CL = 32 * ITEMS;
    constexpr int PP = (CL + NT - 1) / NT;   // positions per thread in the pre-pass / output pass
    constexpr int VT = ElemTraits<T>::kPerVec;
    constexpr bool kAcc = sizeof(T) == 4;
    extern __shared__ __align__(128) unsigned char smem[];
    float* sdl = reinterpret_cast<float*>(smem);        // [CL] activated delta
    float* sdu = sdl + CL;                              // [CL] delta * u
    float* red = sdu + CL;                              // [NW][CL] per-warp partial outputs
    T* sBC = reinterpret_cast<T*>(red + NW * CL);       // [NW][2 slots][B | C][CL]
    float* sA = reinterpret_cast<float*>(sBC + NW * 4 * CL);
    const int N = p.N;
    float2* sCarry = reinterpret_cast<float2*>(sA + ((N + 3) & ~3));   // [N] (decay product since the row start, state)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SPW = (N + NW - 1) / NW;     // states per warp
    const int NWV = N < NW ? N : NW;       // warps that own at least one state
    const int nt = p.nxchunks;
    const int e0 = lane * ITEMS;
    const int64_t nrows = (int64_t)p.batch * p.dim;

    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int b = (int)(row / p.dim), d = (int)(row - (int64_t)b * p.dim), g = d / p.Dg;
        __syncthreads();   // the previous row's shared memory is no longer read
        for (int n = tid; n < N; n += NT) {
            sA[n] = p.A[(int64_t)d * p.A_ds + (int64_t)n * p.A_ns];
            sCarry[n] = make_float2(1.f, 0.f);
        }
        const float Dv = p.D ? p.D[d] : 0.f, bias = p.bias ? p.bias[d] : 0.f;
        const T* gu = reinterpret_cast<const T*>(p.u) + b * p.u_bs + d * p.u_ds;
        const T* gd = reinterpret_cast<const T*>(p.delta) + b * p.dl_bs + d * p.dl_ds;
        const T* gB = reinterpret_cast<const T*>(p.Bm) + b * p.B_bs + g * p.B_gs;
        const T* gC = reinterpret_cast<const T*>(p.Cm) + b * p.C_bs + g * p.C_gs;
        OutT* gout = reinterpret_cast<OutT*>(p.out) + b * p.out_bs + d * p.out_ds;
        float2* gx = p.x ? reinterpret_cast<float2*>(p.x) + row * nt * N : nullptr;

        float ru[PP], rd[PP], Du[PP];
        auto load_ud = [&](int c) {
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int pos = tid + q * NT;
                const int l = c * CL + pos;
                const bool valid = pos < CL && l < p.L;
                ru[q] = valid ? ldg_early(gu + l) : 0.f;
                rd[q] = valid ? ldg_early(gd + l) : 0.f;
            }
        };
        auto prepass = [&](int c) {
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int pos = tid + q * NT;
                if (pos < CL) {
                    float xd = rd[q] + bias;
                    if (p.softplus) xd = softplus_f(xd);
                    if (c * CL + pos >= p.L) xd = 0.f;   // identity beyond the sequence end
                    sdl[pos] = xd;
                    sdu[pos] = xd * ru[q];
                    Du[q] = Dv * ru[q];
                }
            }
        };
        // work item `it` of this warp = (chunk it / SPW, state warp + NW * (it % SPW)); its B / C slices go to slot it & 1.
        // (pc, pj) walk the items in issue order, two items ahead of the one being processed.
        const bool bc_aligned = aligned16(gB, p.B_ns * (int64_t)sizeof(T)) && aligned16(gC, p.C_ns * (int64_t)sizeof(T));
        int pc = 0, pj = 0, pit = 0;
        auto issue = [&]() {
            if (pc < nt) {
                const int n = warp + NW * pj;
                if (n < N) {
                    const int l0 = pc * CL;
                    T* slot = sBC + (size_t)((warp * 2 + (pit & 1)) * 2) * CL;
                    const T* srcB = gB + (int64_t)n * p.B_ns + l0;
                    const T* srcC = gC + (int64_t)n * p.C_ns + l0;
                    if (bc_aligned && l0 + CL <= p.L) {
                        stage_full<T, ITEMS>(slot, srcB, lane);
                        stage_full<T, ITEMS>(slot + CL, srcC, lane);
                    } else {
                        const int len = min(CL, p.L - l0);
                        stage_slice<T>(slot, srcB, len, CL, lane);
                        stage_slice<T>(slot + CL, srcC, len, CL, lane);
                    }
                }
                if (++pj == SPW) {
                    pj = 0;
                    ++pc;
                }
            }
            ++pit;
            cp_async_commit();   // one group per item, empty or not: the wait below counts groups
        };
        load_ud(0);
        prepass(0);
        if (nt > 1) load_ud(1);
        issue();
        issue();
        __syncthreads();

        int it = 0;
        for (int c = 0; c < nt; ++c) {
            const int l0 = c * CL, len = min(CL, p.L - l0);
            float y[ITEMS];
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) y[i] = 0.f;
            for (int j = 0; j < SPW; ++j, ++it) {
                const int n = warp + NW * j;
                cp_async_wait<1>();   // every group but the newest has landed: this item's slices are in shared memory
                __syncwarp();
                if (n < N) {
                    const T* sB = sBC + (size_t)((warp * 2 + (it & 1)) * 2) * CL;
                    const T* sC = sB + CL;
                    const float An = sA[n];
                    float h[ITEMS], rp[ITEMS];
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int v = 0; v < ITEMS / VT; ++v) {
                        float dl[VT], du[VT], Bv[VT];
                        lds_items<float, VT>(sdl + e0 + v * VT, dl);
                        lds_items<float, VT>(sdu + e0 + v * VT, du);
                        lds_items<T, VT>(sB + e0 + v * VT, Bv);
                        if constexpr (kAcc) {
#pragma unroll
                            for (int k = 0; k < VT; k += 2) {
                                float ee[2], bb[2];
                                upk2(decay_m1_2(mul2(pk2(dl[k], dl[k + 1]), splat2(An))), ee[0], ee[1]);
                                upk2(mul2(pk2(du[k], du[k + 1]), pk2(Bv[k], Bv[k + 1])), bb[0], bb[1]);
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    decay_step(ee[q], bb[q], P, V);
                                    h[v * VT + k + q] = V;
                                    rp[v * VT + k + q] = P;
                                }
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < VT; ++k) {
                                decay_step(decay_m1<false>(dl[k] * An), du[k] * Bv[k], P, V);
                                h[v * VT + k] = V;
                                rp[v * VT + k] = P;
                            }
                        }
                    }
                    warp_scan_fwd(P, V, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                    if (lane == 0) {
                        Pe = 1.f;
                        Ve = 0.f;
                    }
                    const float2 carry = sCarry[n];
                    const float seed = fmaf(Pe, carry.y, Ve);
                    const float Pa = __shfl_sync(FULL, P, 31), Va = __shfl_sync(FULL, V, 31);
                    __syncwarp();   // every lane has read the carry
                    if (lane == 0) {
                        const float2 nc = make_float2(carry.x * Pa, fmaf(Pa, carry.y, Va));
                        sCarry[n] = nc;
                        if (gx) gx[(int64_t)c * N + n] = nc;
                    }
#pragma unroll
                    for (int v = 0; v < ITEMS / VT; ++v) {
                        float Cv[VT];
                        lds_items<T, VT>(sC + e0 + v * VT, Cv);
#pragma unroll
                        for (int k = 0; k < VT; ++k) {
                            const int i = v * VT + k;
                            y[i] = fmaf(Cv[k], fmaf(rp[i], seed, h[i]), y[i]);
                        }
                    }
                }
                __syncwarp();    // the slot is free again
                issue();
            }
            if (warp < NWV) sts_items<float, ITEMS>(red + warp * CL + e0, y);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int pos = tid + q * NT;
                if (pos < len) {
                    float s = Du[q];
                    if (NWV == NW) {
#pragma unroll
                        for (int w = 0; w < NW; ++w) s += red[w * CL + pos];
                    } else {
                        for (int w = 0; w < NWV; ++w) s += red[w * CL + pos];
                    }
                    gout[l0 + pos] = ElemTraits<OutT>::from_f(s);
                }
            }
            if (c + 1 < nt) {
                prepass(c + 1);
                if (c + 2 < nt) load_ud(c + 2);
            }
            __syncthreads();
        }
        cp_async_wait<0>();
    }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
template <typename T, typename DT, int ITEMS>
__global__ void __launch_bounds__(kRowsWarps * 32, 2) scan_rows_bwd_kernel(const ScanBwdArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NW = kRowsWarps, NT = NW * 32, CL = 32 * ITEMS;
    constexpr int PP = (CL + NT - 1) / NT;
    constexpr int VT = ElemTraits<T>::kPerVec;
    constexpr bool kAcc = sizeof(T) == 4;
    extern __shared__ __align__(128) unsigned char smem[];
    float* sdl = reinterpret_cast<float*>(smem);        // [CL] activated delta
    float* su = sdl + CL;                               // [CL] u
    float* sdy = su + CL;                               // [CL] dout
    float* redU = sdy + CL;                             // [NW][CL] per-warp partial du
    float* redD = redU + NW * CL;                       // [NW][CL] per-warp partial ddelta
    T* sBC = reinterpret_cast<T*>(redD + NW * CL);      // [NW][2 slots][B | C][CL]
    float* sHin = reinterpret_cast<float*>(sBC + NW * 4 * CL);   // [NW][2 slots] forward state entering the item's chunk
    float* sA = sHin + NW * 2;
    const int N = p.N;
    const int Npad = (N + 3) & ~3;
    float* sR = sA + Npad;      // [N] adjoint entering the chunk from the right
    float* sdA = sR + Npad;     // [N][32] dA of this row, one partial sum per lane of the owning warp (reduced once per row)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SPW = (N + NW - 1) / NW;
    const int NWV = N < NW ? N : NW;
    const int nt = p.nchunks;
    const int e0 = lane * ITEMS;
    const int64_t nrows = (int64_t)p.batch * p.dim;

    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int b = (int)(row / p.dim), d = (int)(row - (int64_t)b * p.dim), g = d / p.Dg;
        __syncthreads();
        for (int n = tid; n < N; n += NT) {
            sA[n] = p.A[(int64_t)d * p.A_ds + (int64_t)n * p.A_ns];
            sR[n] = 0.f;
        }
        for (int i = tid; i < N * 32; i += NT) sdA[i] = 0.f;
        const float Dv = p.D ? p.D[d] : 0.f, bias = p.bias ? p.bias[d] : 0.f;
        const T* gu = reinterpret_cast<const T*>(p.u) + b * p.u_bs + d * p.u_ds;
        const T* gd = reinterpret_cast<const T*>(p.delta) + b * p.dl_bs + d * p.dl_ds;
        const DT* gdy = reinterpret_cast<const DT*>(p.dout) + b * p.do_bs + d * p.do_ds;
        const T* gB = reinterpret_cast<const T*>(p.Bm) + b * p.B_bs + g * p.B_gs;
        const T* gC = reinterpret_cast<const T*>(p.Cm) + b * p.C_bs + g * p.C_gs;
        T* gdu = reinterpret_cast<T*>(p.du) + b * p.du_bs + d * p.du_ds;
        T* gdd = reinterpret_cast<T*>(p.ddelta) + b * p.dd_bs + d * p.dd_ds;
        float* gdB = p.dB + ((int64_t)b * p.G + g) * N * p.L;
        float* gdC = p.dC + ((int64_t)b * p.G + g) * N * p.L;
        const float* gx = p.x ? p.x + row * nt * N * 2 : nullptr;
        const bool vec_red = (p.L & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.dB) | reinterpret_cast<uintptr_t>(p.dC)) & 15) == 0;

        float ru[PP], rd[PP], ry[PP];       // the chunk being loaded
        float ku[PP], kx[PP], ky[PP];       // the chunk being processed (kept for the output pass)
        float dD_acc = 0.f, dbias_acc = 0.f;
        auto load_in = [&](int c) {
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int pos = tid + q * NT;
                const int l = c * CL + pos;
                const bool valid = pos < CL && l < p.L;
                ru[q] = valid ? ldg_early(gu + l) : 0.f;
                rd[q] = valid ? ldg_early(gd + l) : 0.f;
                ry[q] = valid ? ldg_early(gdy + l) : 0.f;
            }
        };
        auto prepass = [&](int c) {
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int pos = tid + q * NT;
                if (pos < CL) {
                    float xd = rd[q] + bias;
                    if (p.softplus) xd = softplus_f(xd);
                    if (c * CL + pos >= p.L) xd = 0.f;
                    sdl[pos] = xd;
                    su[pos] = ru[q];
                    sdy[pos] = ry[q];
                    kx[q] = xd;
                    ku[q] = ru[q];
                    ky[q] = ry[q];
                }
            }
        };
        // work item `it` = (chunk nt - 1 - it / SPW, state warp + NW * (it % SPW)); (pc, pj) walk them two items ahead
        const bool bc_aligned = aligned16(gB, p.B_ns * (int64_t)sizeof(T)) && aligned16(gC, p.C_ns * (int64_t)sizeof(T));
        int pc = nt - 1, pj = 0, pit = 0;
        auto issue = [&]() {
            if (pc >= 0) {
                const int n = warp + NW * pj;
                if (n < N) {
                    const int l0 = pc * CL;
                    T* slot = sBC + (size_t)((warp * 2 + (pit & 1)) * 2) * CL;
                    const T* srcB = gB + (int64_t)n * p.B_ns + l0;
                    const T* srcC = gC + (int64_t)n * p.C_ns + l0;
                    if (bc_aligned && l0 + CL <= p.L) {
                        stage_full<T, ITEMS>(slot, srcB, lane);
                        stage_full<T, ITEMS>(slot + CL, srcC, lane);
                    } else {
                        const int len = min(CL, p.L - l0);
                        stage_slice<T>(slot, srcB, len, CL, lane);
                        stage_slice<T>(slot + CL, srcC, len, CL, lane);
                    }
                    if (lane == 0) {
                        float* hs = sHin + warp * 2 + (pit & 1);
                        if (pc > 0 && gx) cp_async4(hs, gx + ((int64_t)(pc - 1) * N + n) * 2 + 1);
                        else *hs = 0.f;
                    }
                }
                if (++pj == SPW) {
                    pj = 0;
                    --pc;
                }
            }
            ++pit;
            cp_async_commit();
        };
        load_in(nt - 1);
        prepass(nt - 1);
        if (nt > 1) load_in(nt - 2);
        issue();
        issue();
        __syncthreads();

        int it = 0;
        for (int c = nt - 1; c >= 0; --c) {
            const int l0 = c * CL, len = min(CL, p.L - l0);
            // du / ddelta summed over this warp's states accumulate in the warp's own rows of redU / redD (registers are what
            // limits this kernel to two CTAs per SM)
            float* accU = redU + warp * CL + e0;
            float* accD = redD + warp * CL + e0;
            for (int j = 0; j < SPW; ++j, ++it) {
                const int n = warp + NW * j;
                cp_async_wait<1>();
                __syncwarp();
                if (n < N) {
                    const T* sB = sBC + (size_t)((warp * 2 + (it & 1)) * 2) * CL;
                    const T* sC = sB + CL;
                    const float An = sA[n];
                    const float h_in = sHin[warp * 2 + (it & 1)];
                    float a[ITEMS], h[ITEMS];
                    // forward states of the chunk from the carry of the forward pass
                    float P = 1.f, V = 0.f;
                    {
                        float rp[ITEMS];
#pragma unroll
                        for (int v = 0; v < ITEMS / VT; ++v) {
                            float dl[VT], uv[VT], Bv[VT];
                            lds_items<float, VT>(sdl + e0 + v * VT, dl);
                            lds_items<float, VT>(su + e0 + v * VT, uv);
                            lds_items<T, VT>(sB + e0 + v * VT, Bv);
                            if constexpr (kAcc) {
#pragma unroll
                                for (int k = 0; k < VT; k += 2) {
                                    const f32x2 x2 = pk2(dl[k], dl[k + 1]);
                                    float ee[2], bb[2];
                                    upk2(decay_m1_2(mul2(x2, splat2(An))), ee[0], ee[1]);
                                    upk2(mul2(mul2(x2, pk2(uv[k], uv[k + 1])), pk2(Bv[k], Bv[k + 1])), bb[0], bb[1]);
#pragma unroll
                                    for (int q = 0; q < 2; ++q) {
                                        const int i = v * VT + k + q;
                                        a[i] = ee[q];
                                        decay_step(ee[q], bb[q], P, V);
                                        h[i] = V;
                                        rp[i] = P;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < VT; ++k) {
                                    const int i = v * VT + k;
                                    const float ei = decay_m1<false>(dl[k] * An);
                                    const float bi = dl[k] * uv[k] * Bv[k];
                                    a[i] = ei;
                                    decay_step(ei, bi, P, V);
                                    h[i] = V;
                                    rp[i] = P;
                                }
                            }
                        }
                        const float Pth0 = P;
                        warp_scan_fwd(P, V, lane);
                        float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                        if (lane == 0) {
                            Pe = 1.f;
                            Ve = 0.f;
                        }
                        const float seed = fmaf(Pe, h_in, Ve);
#pragma unroll
                        for (int i = 0; i < ITEMS; ++i) h[i] = fmaf(rp[i], seed, h[i]);
                        P = Pth0;   // product of this lane's decays
                    }
                    // adjoint g_t = C_t dout_t + a_{t+1} g_{t+1}. First the lane aggregate with zero incoming adjoint (the decay
                    // product of the lane is P from above), then the warp scan, then the true adjoints together with the outputs:
                    // nothing per position is kept between the two passes.
                    float r = 0.f;
#pragma unroll
                    for (int v = ITEMS / VT - 1; v >= 0; --v) {
                        float Cv[VT], dy[VT];
                        lds_items<T, VT>(sC + e0 + v * VT, Cv);
                        lds_items<float, VT>(sdy + e0 + v * VT, dy);
#pragma unroll
                        for (int k = VT - 1; k >= 0; --k) {
                            const float gz = fmaf(Cv[k], dy[k], r);
                            r = fmaf(a[v * VT + k], gz, gz);
                        }
                    }
                    float Pr = P, Rr = r;
                    warp_scan_rev(Pr, Rr, lane);
                    float Ps = __shfl_down_sync(FULL, Pr, 1), Rs = __shfl_down_sync(FULL, Rr, 1);
                    if (lane == 31) {
                        Ps = 1.f;
                        Rs = 0.f;
                    }
                    const float r_in = sR[n];
                    const float Pa = __shfl_sync(FULL, Pr, 0), Ra = __shfl_sync(FULL, Rr, 0);
                    __syncwarp();
                    if (lane == 0) sR[n] = fmaf(Pa, r_in, Ra);
                    r = fmaf(Ps, r_in, Rs);   // a_{t+1} g_{t+1} entering this lane's last position
                    float dA_acc = 0.f;
                    float* dBn = gdB + (int64_t)n * p.L + l0 + e0;
                    float* dCn = gdC + (int64_t)n * p.L + l0 + e0;
#pragma unroll
                    for (int v = ITEMS / VT - 1; v >= 0; --v) {
                        float dl[VT], uv[VT], Bv[VT], dy[VT], Cv[VT], cB[VT], cC[VT], aU[VT], aD[VT];
                        lds_items<float, VT>(sdl + e0 + v * VT, dl);
                        lds_items<float, VT>(su + e0 + v * VT, uv);
                        lds_items<T, VT>(sB + e0 + v * VT, Bv);
                        lds_items<T, VT>(sC + e0 + v * VT, Cv);
                        lds_items<float, VT>(sdy + e0 + v * VT, dy);
                        if (j > 0) {
                            lds_items<float, VT>(accU + v * VT, aU);
                            lds_items<float, VT>(accD + v * VT, aD);
                        } else {
#pragma unroll
                            for (int k = 0; k < VT; ++k) aU[k] = aD[k] = 0.f;
                        }
#pragma unroll
                        for (int k = VT - 1; k >= 0; --k) {
                            const int i = v * VT + k;
                            const float gt = fmaf(Cv[k], dy[k], r);
                            r = fmaf(a[i], gt, gt);
                            const float bi = dl[k] * uv[k] * Bv[k];
                            const float tt = gt * (h[i] - bi);        // g_t a_t h_{t-1}
                            const float gx_ = gt * dl[k];
                            aU[k] = fmaf(gx_, Bv[k], aU[k]);
                            aD[k] += fmaf(gt * uv[k], Bv[k], An * tt);
                            dA_acc = fmaf(dl[k], tt, dA_acc);
                            cB[k] = gx_ * uv[k];
                            cC[k] = dy[k] * h[i];
                        }
                        sts_items<float, VT>(accU + v * VT, aU);
                        sts_items<float, VT>(accD + v * VT, aD);
                        if (vec_red && len == CL) {
#pragma unroll
                            for (int k = 0; k < VT; k += 4) {
                                red_add_v4(dBn + v * VT + k, cB[k], cB[k + 1], cB[k + 2], cB[k + 3]);
                                red_add_v4(dCn + v * VT + k, cC[k], cC[k + 1], cC[k + 2], cC[k + 3]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < VT; ++k) {
                                if (e0 + v * VT + k < len) {
                                    atomicAdd(dBn + v * VT + k, cB[k]);
                                    atomicAdd(dCn + v * VT + k, cC[k]);
                                }
                            }
                        }
                    }
                    sdA[n * 32 + lane] += dA_acc;
                }
                __syncwarp();
                issue();
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int pos = tid + q * NT;
                if (pos < len) {
                    float sU = Dv * ky[q], sD = 0.f;
                    if (NWV == NW) {
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            sU += redU[w * CL + pos];
                            sD += redD[w * CL + pos];
                        }
                    } else {
                        for (int w = 0; w < NWV; ++w) {
                            sU += redU[w * CL + pos];
                            sD += redD[w * CL + pos];
                        }
                    }
                    // d softplus = sigmoid(raw) = 1 - exp(-softplus(raw)) = -expm1(-x) (bwd_kernel_oflex.cuh:250-255)
                    if (p.softplus) sD *= -decay_m1<true>(-kx[q]);
                    gdu[l0 + pos] = ElemTraits<T>::from_f(sU);
                    gdd[l0 + pos] = ElemTraits<T>::from_f(sD);
                    dD_acc = fmaf(ky[q], ku[q], dD_acc);
                    dbias_acc += sD;
                }
            }
            if (c > 0) {
                prepass(c - 1);
                if (c > 1) load_in(c - 2);
            }
            __syncthreads();
        }
        cp_async_wait<0>();
        for (int n = warp; n < N; n += NW) {   // warp w owns states w, w + NW, ...: its own partial sums, no barrier needed
            float v = sdA[n * 32 + lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (lane == 0) atomicAdd(p.dA + (int64_t)d * N + n, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dD_acc += __shfl_xor_sync(FULL, dD_acc, o);
            dbias_acc += __shfl_xor_sync(FULL, dbias_acc, o);
        }
        if (lane == 0) {
            if (p.dD) atomicAdd(p.dD + d, dD_acc);
            if (p.dbias) atomicAdd(p.dbias + d, dbias_acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// The row kernels pay off when there are rows enough to fill the machine; below that the look-back kernels (which split L
// across CTAs) win. dstate > kMaxDstate only runs here. BEM_SCAN_ROWS=0 / 1 forces the choice (A/B measurements, tests).
bool scan_rows_preferred(int batch, int dim, int N, int sm_count) {
    if (N < 2) return false;
    if (N > kMaxDstate) return true;
    if (const char* ev = getenv("BEM_SCAN_ROWS")) return atoi(ev) != 0;
    return (int64_t)batch * dim >= sm_count;
}

template <typename K>
static int set_smem(K kernel, int smem_bytes, int* cache) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cache[dev] < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        cache[dev] = smem_bytes;
    }
    return 0;
}

template <typename T, typename OutT, int ITEMS>
static int launch_rows_fwd(const ScanFwdArgs& a, int sm_count, cudaStream_t stream) {
    constexpr int CL = 32 * ITEMS, NW = kRowsWarps;
    auto kernel = scan_rows_fwd_kernel<T, OutT, ITEMS>;
    const int Npad = (a.N + 3) & ~3;
    const int smem_bytes = (2 + NW) * CL * 4 + NW * 4 * CL * (int)sizeof(T) + Npad * 4 + a.N * 8 + 16;
    static int cache[64] = {0};
    if (int rc = set_smem(kernel, smem_bytes, cache)) return rc;
    const int64_t rows = (int64_t)a.batch * a.dim;
    (void)sm_count;
    const int grid = (int)(rows < 0x7fffffffLL ? rows : 0x7fffffffLL);   // one CTA per row; the row loop only matters beyond 2^31 rows
    launch_pdl(kernel, dim3(grid), dim3(NW * 32), smem_bytes, stream, a);
    return (int)cudaGetLastError();
}

template <typename T, typename DT, int ITEMS>
static int launch_rows_bwd(const ScanBwdArgs& a, int sm_count, cudaStream_t stream) {
    constexpr int CL = 32 * ITEMS, NW = kRowsWarps;
    auto kernel = scan_rows_bwd_kernel<T, DT, ITEMS>;
    const int Npad = (a.N + 3) & ~3;
    const int smem_bytes = (3 + 2 * NW) * CL * 4 + NW * 4 * CL * (int)sizeof(T) + NW * 2 * 4 + 2 * Npad * 4 + a.N * 32 * 4 + 16;
    static int cache[64] = {0};
    if (int rc = set_smem(kernel, smem_bytes, cache)) return rc;
    const int64_t rows = (int64_t)a.batch * a.dim;
    (void)sm_count;
    const int grid = (int)(rows < 0x7fffffffLL ? rows : 0x7fffffffLL);   // one CTA per row
    launch_pdl(kernel, dim3(grid), dim3(NW * 32), smem_bytes, stream, a);
    return (int)cudaGetLastError();
}

int scan_rows_fwd_dispatch(const ScanFwdArgs& a, int dtype, int out_dtype, int sm_count, cudaStream_t stream) {
    if (a.N < 2 || a.R > 0) return BEM_ERR_UNSUPPORTED;
    if (dtype == BEM_F32) return launch_rows_fwd<float, float, kItemsF32>(a, sm_count, stream);
    if (dtype == BEM_F16) {
        if (out_dtype == BEM_F32) return launch_rows_fwd<__half, float, kItems16>(a, sm_count, stream);
        return launch_rows_fwd<__half, __half, kItems16>(a, sm_count, stream);
    }
    if (dtype == BEM_BF16) {
        if (out_dtype == BEM_F32) return launch_rows_fwd<__nv_bfloat16, float, kItems16>(a, sm_count, stream);
        return launch_rows_fwd<__nv_bfloat16, __nv_bfloat16, kItems16>(a, sm_count, stream);
    }
    return BEM_ERR_BAD_ARG;
}

int scan_rows_bwd_dispatch(const ScanBwdArgs& a, int dtype, int dout_dtype, int sm_count, cudaStream_t stream) {
    if (a.N < 2) return BEM_ERR_UNSUPPORTED;
    if (dtype == BEM_F32) return launch_rows_bwd<float, float, kItemsF32>(a, sm_count, stream);
    if (dtype == BEM_F16) {
        if (dout_dtype == BEM_F32) return launch_rows_bwd<__half, float, kItems16>(a, sm_count, stream);
        return launch_rows_bwd<__half, __half, kItems16>(a, sm_count, stream);
    }
    if (dtype == BEM_BF16) {
        if (dout_dtype == BEM_F32) return launch_rows_bwd<__nv_bfloat16, float, kItems16>(a, sm_count, stream);
        return launch_rows_bwd<__nv_bfloat16, __nv_bfloat16, kItems16>(a, sm_count, stream);
    }
    return BEM_ERR_BAD_ARG;
}

}  // namespace bem
