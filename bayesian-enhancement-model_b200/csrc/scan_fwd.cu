// scan_fwd.cu — selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cusoflex/
// selective_scan_fwd_kernel_oflex.cuh:67-180) and its launcher (:182-211). Same math:
//   delta = softplus(delta + bias); a = exp(delta * A_n); b = delta * u * B_n; h = a h + b; y = D u + sum_n C_n h
// Different organisation:
//   * persistent CTAs: 1 producer warp + NW consumer warps; a tile is NW channel rows of one B/C group x one chunk
//     of CL = 32*ITEMS positions. The producer stages u/delta rows and the shared B/C chunk into a ring of shared
//     memory stages with TMA bulk copies (mbarrier transaction counts); B/C are fetched once per tile, not once per
//     channel row as in the reference (fwd_kernel_oflex.cuh:137-140)
//   * each consumer warp owns one row: lane-local sequential scan over ITEMS consecutive positions, warp-shuffle
//     scan across lanes, decoupled look-back across chunks -> L is parallel across CTAs
//   * y is written back through shared memory with a TMA bulk store
#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

template <typename T, typename OutT, int ITEMS, int NW, bool N1>
__global__ void __launch_bounds__((NW + 1) * 32) scan_fwd_kernel(const ScanFwdArgs p) {
    constexpr int CL = 32 * ITEMS;
    constexpr bool kAcc = sizeof(T) == 4;   // fp32 inputs: <= 1 ulp decay factors (scan_common.cuh decay_m1)
    constexpr int ROW_SLOT = 2 * CL * (int)sizeof(T);   // [u chunk | delta chunk]; y (OutT) is written over it
    static_assert(CL * sizeof(OutT) <= (size_t)ROW_SLOT, "output overlay must fit the row slot");
    extern __shared__ __align__(128) unsigned char smem[];

    const int N = N1 ? 1 : p.N;
    const int S = p.stages;
    const int bc_bytes = N * CL * (int)sizeof(T);
    const int stage_bytes = NW * ROW_SLOT + 2 * bc_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;
    int* tile_slot = reinterpret_cast<int*>(empty + S);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int RT = p.RT;
    const int GRB = p.G * p.RB;

    auto decode = [&](int t, TileCoord& tc) {
        tc.c = t / RT;
        const int r = t - tc.c * RT;
        tc.b = r / GRB;
        const int rem = r - tc.b * GRB;
        tc.g = rem / p.RB;
        const int rb = rem - tc.g * p.RB;
        tc.row0 = rb * NW;
        tc.nrows = min(NW, p.Dg - tc.row0);
    };

    if (warp == NW) {
        // ======================================= producer warp =======================================
        for (uint32_t it = 0;; ++it) {
            const int s = it % S;
            const uint32_t use = it / S;
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(p.ticket, 1u);
            t = __shfl_sync(FULL, t, 0);
            if (t >= (unsigned)p.total_tiles) {
                if (lane == 0) {
                    tile_slot[s] = -1;
                    mbar_arrive(&full[s]);
                }
                break;
            }
            TileCoord tc;
            decode((int)t, tc);
            const int l0 = tc.c * CL;
            const int len = min(CL, p.L - l0);
            unsigned char* st = smem + (size_t)s * stage_bytes;
            // jobs: [0, nrows) u rows, [nrows, 2 nrows) delta rows, then N B rows, N C rows
            const int njobs = 2 * tc.nrows + 2 * N;
            uint32_t my_bytes = 0;
            for (int j = lane; j < njobs; j += 32) {
                const T* src;
                T* dst;
                if (j < 2 * tc.nrows) {
                    const int isd = j >= tc.nrows;
                    const int r = isd ? j - tc.nrows : j;
                    const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + r;
                    src = isd ? reinterpret_cast<const T*>(p.delta) + tc.b * p.dl_bs + d * p.dl_ds + l0
                              : reinterpret_cast<const T*>(p.u) + tc.b * p.u_bs + d * p.u_ds + l0;
                    dst = reinterpret_cast<T*>(st + r * ROW_SLOT) + (isd ? CL : 0);
                } else {
                    const int k = j - 2 * tc.nrows;
                    const int isc = k >= N;
                    const int n = isc ? k - N : k;
                    src = isc ? reinterpret_cast<const T*>(p.Cm) + tc.b * p.C_bs + tc.g * p.C_gs + n * p.C_ns + l0
                              : reinterpret_cast<const T*>(p.Bm) + tc.b * p.B_bs + tc.g * p.B_gs + n * p.B_ns + l0;
                    dst = reinterpret_cast<T*>(st + NW * ROW_SLOT + (isc ? bc_bytes : 0)) + n * CL;
                }
                const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
                const uint32_t vec_bytes = aligned ? ((uint32_t)(len * sizeof(T)) & ~15u) : 0u;
                // ragged tail (or an unaligned row): plain loads by this lane
                for (int e = vec_bytes / sizeof(T); e < len; ++e) dst[e] = src[e];
                my_bytes += vec_bytes;
            }
            uint32_t tot = my_bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            __syncwarp();
            if (lane == 0) {
                tile_slot[s] = (int)t;
                if (tot > 0) mbar_arrive_expect_tx(&full[s], tot);
                else mbar_arrive(&full[s]);
            }
            __syncwarp();
            for (int j = lane; j < njobs; j += 32) {
                const T* src;
                T* dst;
                if (j < 2 * tc.nrows) {
                    const int isd = j >= tc.nrows;
                    const int r = isd ? j - tc.nrows : j;
                    const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + r;
                    src = isd ? reinterpret_cast<const T*>(p.delta) + tc.b * p.dl_bs + d * p.dl_ds + l0
                              : reinterpret_cast<const T*>(p.u) + tc.b * p.u_bs + d * p.u_ds + l0;
                    dst = reinterpret_cast<T*>(st + r * ROW_SLOT) + (isd ? CL : 0);
                } else {
                    const int k = j - 2 * tc.nrows;
                    const int isc = k >= N;
                    const int n = isc ? k - N : k;
                    src = isc ? reinterpret_cast<const T*>(p.Cm) + tc.b * p.C_bs + tc.g * p.C_gs + n * p.C_ns + l0
                              : reinterpret_cast<const T*>(p.Bm) + tc.b * p.B_bs + tc.g * p.B_gs + n * p.B_ns + l0;
                    dst = reinterpret_cast<T*>(st + NW * ROW_SLOT + (isc ? bc_bytes : 0)) + n * CL;
                }
                const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
                const uint32_t vec_bytes = aligned ? ((uint32_t)(len * sizeof(T)) & ~15u) : 0u;
                if (vec_bytes) bulk_g2s(dst, src, vec_bytes, &full[s]);
            }
        }
        return;
    }

    // ========================================= consumer warps =========================================
    int pend_stage = -1;   // stage whose bulk store has been issued but not yet drained
    for (uint32_t it = 0;; ++it) {
        const int s = it % S;
        mbar_wait(&full[s], (it / S) & 1, p.err);
        const int t = tile_slot[s];
        if (t < 0) break;
        TileCoord tc;
        decode(t, tc);
        const bool active = warp < tc.nrows;
        unsigned char* st = smem + (size_t)s * stage_bytes;
        bool drained = false;
        auto drain_prev = [&]() {   // release the previous stage once its y store has left shared memory
            if (!drained && pend_stage >= 0 && lane == 0) {
                bulk_wait_read<0>();
                mbar_arrive(&empty[pend_stage]);
            }
            drained = true;
        };
        if (active) {
            const int c = tc.c;
            const int l0 = c * CL;
            const int len = min(CL, p.L - l0);
            const bool partial = len < CL;
            const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + warp;
            const int64_t row = (int64_t)tc.b * p.dim + d;
            const T* su = reinterpret_cast<const T*>(st + warp * ROW_SLOT);
            const T* sB = reinterpret_cast<const T*>(st + NW * ROW_SLOT);
            const T* sC = reinterpret_cast<const T*>(st + NW * ROW_SLOT + bc_bytes);
            const int e0 = lane * ITEMS;

            float uv[ITEMS], dl[ITEMS];
            lds_items<T, ITEMS>(su + e0, uv);
            lds_items<T, ITEMS>(su + CL + e0, dl);
            const float bias = p.bias ? p.bias[d] : 0.f;
            const float Dv = p.D ? p.D[d] : 0.f;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                float x = dl[i] + bias;
                if (p.softplus) x = softplus_f(x);
                dl[i] = x;
            }
            float y[ITEMS];

            if constexpr (N1) {
                const float A2 = p.A[d * p.A_ds];
                float cumA[ITEMS], hloc[ITEMS];
                {
                    float Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + e0, Bv);
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        float e = decay_m1<kAcc>(dl[i] * A2);
                        float b = dl[i] * uv[i] * Bv[i];
                        if (partial && e0 + i >= len) {   // identity padding so the carried state stays exact
                            e = 0.f;
                            b = 0.f;
                        }
                        decay_step(e, b, P, V);
                        hloc[i] = V;
                        cumA[i] = P;
                    }
                    warp_scan_fwd(P, V, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                    if (lane == 0) {
                        Pe = 1.f;
                        Ve = 0.f;
                    }
                    const float Pa = __shfl_sync(FULL, P, 31), Va = __shfl_sync(FULL, V, 31);
                    uint4* drow = p.desc + row * p.nchunks;
                    float Pp = 1.f, hp = 0.f;
                    if (c > 0) {
                        if (lane == 0 && c + 1 < p.nchunks) st_desc(drow + c, Pa, Va, DESC_AGGREGATE);
                        drain_prev();
                        const float2 pre = lookback(drow, 1, c, p.nchunks, -1, lane, p.err);
                        Pp = pre.x;
                        hp = pre.y;
                    }
                    const float Pi = Pp * Pa, hi = fmaf(Pa, hp, Va);
                    if (lane == 0) {
                        if (c + 1 < p.nchunks) st_desc(drow + c, Pi, hi, DESC_INCLUSIVE);
                        if (p.x) {
                            float2* xr = reinterpret_cast<float2*>(p.x) + row * p.nchunks + c;
                            *xr = make_float2(Pi, hi);
                        }
                    }
                    const float seed = fmaf(Pe, hp, Ve);
                    float Cv[ITEMS];
                    lds_items<T, ITEMS>(sC + e0, Cv);
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const float h = fmaf(cumA[i], seed, hloc[i]);
                        y[i] = fmaf(Cv[i], h, Dv * uv[i]);
                    }
                }
            } else {
                // ---------------- general dstate: aggregates first, look-back, then the seeded pass ----------------
                float du[ITEMS];
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    du[i] = dl[i] * uv[i];
                    y[i] = Dv * uv[i];
                }
                float aggP = 1.f, aggV = 0.f;   // lane n keeps the chunk aggregate of state n
                for (int n = 0; n < N; ++n) {
                    const float A2 = p.A[d * p.A_ds + n * p.A_ns];
                    float Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + n * CL + e0, Bv);
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        float e = decay_m1<kAcc>(dl[i] * A2);
                        float b = du[i] * Bv[i];
                        if (partial && e0 + i >= len) {
                            e = 0.f;
                            b = 0.f;
                        }
                        decay_step(e, b, P, V);
                    }
                    warp_scan_fwd(P, V, lane);
                    const float Pa = __shfl_sync(FULL, P, 31), Va = __shfl_sync(FULL, V, 31);
                    if (lane == n) {
                        aggP = Pa;
                        aggV = Va;
                    }
                }
                uint4* drow = p.desc + (row * p.nchunks) * N;   // [chunk][n]
                float preP = 1.f, preV = 0.f;                   // lane n: composition of chunks < c for state n
                if (c > 0) {
                    if (lane < N && c + 1 < p.nchunks) st_desc(drow + (int64_t)c * N + lane, aggP, aggV, DESC_AGGREGATE);
                    drain_prev();
                    for (int n = 0; n < N; ++n) {
                        const float2 pre = lookback(drow + n, N, c, p.nchunks, -1, lane, p.err);
                        if (lane == n) {
                            preP = pre.x;
                            preV = pre.y;
                        }
                    }
                }
                if (lane < N) {
                    const float Pi = preP * aggP, hi = fmaf(aggP, preV, aggV);
                    if (c + 1 < p.nchunks) st_desc(drow + (int64_t)c * N + lane, Pi, hi, DESC_INCLUSIVE);
                    if (p.x) {
                        float2* xr = reinterpret_cast<float2*>(p.x) + (row * p.nchunks + c) * N + lane;
                        *xr = make_float2(Pi, hi);
                    }
                }
                for (int n = 0; n < N; ++n) {
                    const float A2 = p.A[d * p.A_ds + n * p.A_ns];
                    float Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + n * CL + e0, Bv);
                    float cumA[ITEMS], hloc[ITEMS];
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        float e = decay_m1<kAcc>(dl[i] * A2);
                        float b = du[i] * Bv[i];
                        if (partial && e0 + i >= len) {
                            e = 0.f;
                            b = 0.f;
                        }
                        decay_step(e, b, P, V);
                        hloc[i] = V;
                        cumA[i] = P;
                    }
                    warp_scan_fwd(P, V, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                    if (lane == 0) {
                        Pe = 1.f;
                        Ve = 0.f;
                    }
                    const float hp = __shfl_sync(FULL, preV, n);
                    const float seed = fmaf(Pe, hp, Ve);
                    float Cv[ITEMS];
                    lds_items<T, ITEMS>(sC + n * CL + e0, Cv);
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) y[i] = fmaf(Cv[i], fmaf(cumA[i], seed, hloc[i]), y[i]);
                }
            }

            // ---------------- write y: shared memory overlay + TMA bulk store, scalar tail / unaligned rows ----------------
            drain_prev();
            OutT* gout = reinterpret_cast<OutT*>(p.out) + tc.b * p.out_bs + d * p.out_ds + l0;
            const bool aligned = (reinterpret_cast<uintptr_t>(gout) & 15) == 0;
            const uint32_t vec_bytes = aligned ? ((uint32_t)(len * sizeof(OutT)) & ~15u) : 0u;
            const int vec_elems = vec_bytes / sizeof(OutT);
            __syncwarp();   // every lane has finished reading u/delta of this row before y overwrites the slot
            OutT* sy = reinterpret_cast<OutT*>(st + warp * ROW_SLOT);
            sts_items<OutT, ITEMS>(sy + e0, y);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && vec_bytes) bulk_s2g(gout, sy, vec_bytes);
            if (vec_elems < len) {
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const int e = e0 + i;
                    if (e >= vec_elems && e < len) gout[e] = ElemTraits<OutT>::from_f(y[i]);
                }
            }
        } else {
            drain_prev();
        }
        if (lane == 0) bulk_commit();   // one (possibly empty) group per tile keeps the accounting uniform
        pend_stage = s;
    }
    if (lane == 0) bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename T, typename OutT, int ITEMS, bool N1>
static int launch_fwd(ScanFwdArgs a, int sm_count, cudaStream_t stream) {
    constexpr int NW = kScanWarps;
    constexpr int CL = 32 * ITEMS;
    auto kernel = scan_fwd_kernel<T, OutT, ITEMS, NW, N1>;
    const int stage_bytes = NW * 2 * CL * (int)sizeof(T) + 2 * a.N * CL * (int)sizeof(T);
    // two resident CTAs per SM when three stages fit in half of the shared memory, else one CTA with a deeper ring
    const int budget2 = (227 * 1024) / 2 - 1024;
    int stages, ctas_per_sm;
    if (3 * stage_bytes + 256 <= budget2) {
        stages = min(4, (budget2 - 256) / stage_bytes);
        ctas_per_sm = 2;
    } else {
        stages = min(4, (227 * 1024 - 256) / stage_bytes);
        ctas_per_sm = 1;
        if (stages < 2) return BEM_ERR_UNSUPPORTED;
    }
    a.stages = stages;
    const int smem_bytes = stages * stage_bytes + stages * 2 * 8 + stages * 4 + 64;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    const int grid = min(a.total_tiles, sm_count * ctas_per_sm);
    kernel<<<grid, (NW + 1) * 32, smem_bytes, stream>>>(a);
    return (int)cudaGetLastError();
}

template <typename T, typename OutT, int ITEMS>
static int launch_fwd_n(const ScanFwdArgs& a, int sm_count, cudaStream_t stream) {
    if (a.N == 1) return launch_fwd<T, OutT, ITEMS, true>(a, sm_count, stream);
    return launch_fwd<T, OutT, ITEMS, false>(a, sm_count, stream);
}

int scan_fwd_dispatch(const ScanFwdArgs& a, int dtype, int out_dtype, int sm_count, cudaStream_t stream) {
    if (dtype == BEM_F32) return launch_fwd_n<float, float, kItemsF32>(a, sm_count, stream);
    if (dtype == BEM_F16) {
        if (out_dtype == BEM_F32) return launch_fwd_n<__half, float, kItems16>(a, sm_count, stream);
        return launch_fwd_n<__half, __half, kItems16>(a, sm_count, stream);
    }
    if (dtype == BEM_BF16) {
        if (out_dtype == BEM_F32) return launch_fwd_n<__nv_bfloat16, float, kItems16>(a, sm_count, stream);
        return launch_fwd_n<__nv_bfloat16, __nv_bfloat16, kItems16>(a, sm_count, stream);
    }
    return BEM_ERR_BAD_ARG;
}

}  // namespace bem
