// scan_fwd.cu — selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cusoflex/
// selective_scan_fwd_kernel_oflex.cuh:67-180) and its launcher (:182-211). Same math:
//   delta = softplus(delta + bias); a = exp(delta * A_n); b = delta * u * B_n; h = a h + b; y = D u + sum_n C_n h
// Different organisation:
//   * persistent CTAs: 1 producer warp + NW consumer warps; a tile is NW channel rows of one B/C group x CL = 32*ITEMS
//     positions. The producer stages u/delta rows and the shared B/C chunk into a ring of shared-memory stages with
//     TMA bulk copies (mbarrier transaction counts), plus the tile coordinates and the per-row scalars (A, D, bias),
//     so consumers do no integer division and no dependent global load. B/C are fetched once per tile, not once per
//     channel row as in the reference (fwd_kernel_oflex.cuh:137-140)
//   * each consumer warp owns one row: lane-local sequential scan over ITEMS consecutive positions, warp-shuffle
//     scan across lanes, deterministic decoupled look-back across tiles -> L is parallel across CTAs
//   * y is written straight from registers (each lane owns whole 32-byte sectors); the stage is released before that
//   * carries `x` are emitted every kCarry positions (one or two per tile), so the backward kernel keeps its own tiling
#include <cstdlib>
#include <type_traits>

#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

// stage timeline of CTA 0 (tools/trace_scan.py, env BEM_SCAN_TRACE=1): (tag, arg, SM clock) records, one region per traced
// warp, plain stores — nothing on the critical path waits for them
constexpr int STRACE_ROLES = 4, STRACE_PER = 2048;
__device__ uint4 g_scan_trace[STRACE_ROLES * STRACE_PER];
struct ScanTracer {
    uint32_t n = 0;
    __device__ __forceinline__ void operator()(int on, int role, uint32_t tag, uint32_t arg) {
        if (on && blockIdx.x == 0 && n < STRACE_PER) g_scan_trace[role * STRACE_PER + n++] = make_uint4(tag, arg, (uint32_t)clock64(), 1u);
    }
};

// RANK: fused dt_proj rank. 0 = `delta` given per channel row; > 0 = compile-time rank (the BEM ranks 3 and 5: unrolled, weights
// in registers); -1 = rank read from the arguments (any rank <= kMaxDtRank, runtime loop).
template <typename T, typename OutT, int ITEMS, int NW, bool N1, int RANK = 0>
__global__ void __launch_bounds__((NW + 1) * 32, N1 ? 2 : 1) scan_fwd_kernel(const ScanFwdArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr bool FUSED = RANK != 0;
    constexpr int CL = 32 * ITEMS;
    constexpr bool kAcc = sizeof(T) == 4;   // fp32 inputs: full-precision decay rate (scan_common.cuh decay_m1)
    constexpr int XC = sizeof(T) == 4 ? kCarryF32 : kCarry16;   // positions per carry of `x`
    static_assert(CL % XC == 0, "tile must hold a whole number of carry chunks");
    constexpr int CPT = CL / XC;            // carries per tile
    constexpr int ROW_SLOT = 2 * CL * (int)sizeof(T);   // [u chunk | delta chunk]
    constexpr int V = ElemTraits<T>::kPerVec;
    extern __shared__ __align__(128) unsigned char smem[];

    const int N = N1 ? 1 : p.N;
    const int S = p.stages;
    const int bc_bytes = N * CL * (int)sizeof(T);
    // fused dt_proj (p.R > 0): `delta` is the low-rank dt of the group, (B, G, R, L); the R rows of a tile take the delta
    // halves of row slots 0..R-1 and every channel row forms delta = sum_r W[d][r] * dt[r] from its R weights (scalars)
    const int R = RANK > 0 ? RANK : (RANK < 0 ? p.R : 0);
    const int NSC = N + 2 + R;                                             // scalars per row: A[N], D, bias, W_dt[R]
    const int hdr_bytes = 128 + ((NW * NSC * 4 + 127) / 128) * 128;       // TileCoord | per-row scalars [NW][NSC]
    const int stage_bytes = hdr_bytes + NW * ROW_SLOT + 2 * bc_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    ScanTracer tr;
    const int trace_on = p.trace;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp == NW) {
        // ======================================= producer warp =======================================
        // This tile's per-row scalars are requested before the producer blocks on the stage's empty barrier, and the
        // next ticket while this tile's copies are being issued, so their global-memory latency overlaps useful work.
        const int RT = p.RT;
        const int GRB = p.G * p.RB;
        constexpr int kMaxSc = (NW * (kMaxDstate + 2) + 31) / 32;   // scalars per lane
        // launch epoch of the descriptors: read before this CTA's first draw (the last drawer of the launch bumps it)
        const uint32_t epoch = *reinterpret_cast<volatile unsigned int*>(p.ticket + 2);
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(p.ticket, 1u);
        t = __shfl_sync(FULL, t, 0);
        int s = 0;
        uint32_t use = 0;
        while (true) {
            unsigned char* st = smem + (size_t)s * stage_bytes;
            TileCoord* hdr = reinterpret_cast<TileCoord*>(st);
            if (t >= (unsigned)p.total_tiles) {
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
                if (lane == 0) {
                    hdr->nrows = -1;
                    mbar_arrive(&full[s]);
                    // every CTA ends with exactly one failing draw: the last of them re-arms the workspace for the next
                    // launch (ticket back to zero, new descriptor epoch) — nobody draws after it
                    if (t == (unsigned)p.total_tiles + gridDim.x - 1) {
                        p.ticket[2] = (epoch + 1) & 0x3fffffffu;
                        p.ticket[0] = 0u;
                    }
                }
                break;
            }
            TileCoord tc;
            tc.c = (int)t / RT;
            const int r = (int)t - tc.c * RT;
            tc.b = r / GRB;
            const int rem = r - tc.b * GRB;
            tc.g = rem / p.RB;
            tc.row0 = (rem - tc.g * p.RB) * NW;
            tc.nrows = min(NW, p.Dg - tc.row0);
            const int l0 = tc.c * CL;
            tc.len = min(CL, p.L - l0);
            tc.aux0 = tc.aux1 = 0;
            tc.epoch = epoch;
            const int len = tc.len;
            // per-row scalars A[0..N), D, bias: loads issued now, stored after the slot is free
            float scv[kMaxSc];
#pragma unroll
            for (int q = 0; q < kMaxSc; ++q) {
                const int i = lane + 32 * q;
                float v = 0.f;
                if (i < tc.nrows * NSC) {
                    const int rr = i / NSC, k = i - rr * NSC;
                    const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + rr;
                    if (k < N) v = p.A[d * p.A_ds + k * p.A_ns];
                    else if (k == N) v = p.D ? p.D[d] : 0.f;
                    else if (k == N + 1) v = p.bias ? p.bias[d] : 0.f;
                    else v = p.dt_w[d * R + (k - N - 2)];
                }
                scv[q] = v;
            }
            if (lane == 0) tr(trace_on, 0, 1, t);
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
            if (lane == 0) tr(trace_on, 0, 2, t);
            // the next ticket is drawn only once this slot is free: tickets held ahead of time would sit in this CTA's
            // queue while other CTAs' look-backs wait on them
            unsigned int t_next = 0;
            if (lane == 0) t_next = atomicAdd(p.ticket, 1u);
            if (lane == 0) *hdr = tc;
            float* sc = reinterpret_cast<float*>(st + 128);
#pragma unroll
            for (int q = 0; q < kMaxSc; ++q) {
                const int i = lane + 32 * q;
                if (i < tc.nrows * NSC) sc[i] = scv[q];
            }
            if (lane == 0) tr(trace_on, 0, 6, t);
            unsigned char* rows = st + hdr_bytes;
            // jobs: [0, nrows) u rows, then nd delta rows (one per channel row, or the group's R low-rank rows), N B rows, N C rows
            const int nd = R > 0 ? R : tc.nrows;
            const int njobs = tc.nrows + nd + 2 * N;
            uint32_t my_bytes = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int j = lane; j < njobs; j += 32) {
                    const T* src;
                    T* dst;
                    if (j < tc.nrows + nd) {
                        const int isd = j >= tc.nrows;
                        const int rr = isd ? j - tc.nrows : j;
                        const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + rr;
                        if (!isd) src = reinterpret_cast<const T*>(p.u) + tc.b * p.u_bs + d * p.u_ds + l0;
                        else if (R > 0) src = reinterpret_cast<const T*>(p.delta) + tc.b * p.dl_bs + tc.g * p.dl_gs + rr * p.dl_ds + l0;
                        else src = reinterpret_cast<const T*>(p.delta) + tc.b * p.dl_bs + d * p.dl_ds + l0;
                        dst = reinterpret_cast<T*>(rows + rr * ROW_SLOT) + (isd ? CL : 0);
                    } else {
                        const int k = j - tc.nrows - nd;
                        const int isc = k >= N;
                        const int n = isc ? k - N : k;
                        src = isc ? reinterpret_cast<const T*>(p.Cm) + tc.b * p.C_bs + tc.g * p.C_gs + n * p.C_ns + l0
                                  : reinterpret_cast<const T*>(p.Bm) + tc.b * p.B_bs + tc.g * p.B_gs + n * p.B_ns + l0;
                        dst = reinterpret_cast<T*>(rows + NW * ROW_SLOT + (isc ? bc_bytes : 0)) + n * CL;
                    }
                    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
                    const uint32_t vec_bytes = aligned ? ((uint32_t)(len * sizeof(T)) & ~15u) : 0u;
                    if (pass == 0) {
                        // ragged tail (or an unaligned row): plain loads by this lane
                        for (int e = vec_bytes / sizeof(T); e < len; ++e) dst[e] = src[e];
                        my_bytes += vec_bytes;
                    } else if (vec_bytes) {
                        bulk_g2s(dst, src, vec_bytes, &full[s]);
                    }
                }
                if (pass == 0) {
                    if (lane == 0) tr(trace_on, 0, 7, t);
                    uint32_t tot = my_bytes;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
                    __syncwarp();
                    if (lane == 0) {
                        if (tot > 0) mbar_arrive_expect_tx(&full[s], tot);
                        else mbar_arrive(&full[s]);
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) tr(trace_on, 0, 3, t);
            t = __shfl_sync(FULL, t_next, 0);
            if (++s == S) {
                s = 0;
                ++use;
            }
        }
        return;
    }

    // ========================================= consumer warps =========================================
    const int nt = p.nchunks;   // tiles per row
    int s = -1;
    uint32_t phase = 1;
    while (true) {
        if (++s == S) s = 0;
        if (s == 0) phase ^= 1;
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 10, 0);
        mbar_wait(&full[s], phase, p.err);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        const TileCoord tc = *reinterpret_cast<const TileCoord*>(st);
        const uint32_t ep = tc.epoch;
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 11, tc.c);
        if (tc.nrows < 0) break;
        const bool active = warp < tc.nrows;
        if (active) {
            const int c = tc.c;
            const int l0 = c * CL;
            const int len = tc.len;
            const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + warp;
            const int64_t row = (int64_t)tc.b * p.dim + d;
            const float* sc = reinterpret_cast<const float*>(st + 128) + warp * NSC;
            unsigned char* rows = st + hdr_bytes;
            const T* su = reinterpret_cast<const T*>(rows + warp * ROW_SLOT);
            const T* sB = reinterpret_cast<const T*>(rows + NW * ROW_SLOT);
            const T* sC = reinterpret_cast<const T*>(rows + NW * ROW_SLOT + bc_bytes);
            const int e0 = lane * ITEMS;
            const float Dv = sc[N], bias = sc[N + 1];
            const LookbackPlan plan = lookback_plan(c, nt);
            float y[ITEMS];

            if constexpr (N1) {
                uint4* aggrow = p.desc + row * nt;
                uint4* inclrow = p.desc_incl + row * nt;
                const uint4* lb_addr = p.lb_dynamic ? nullptr : lookback_addr(aggrow, inclrow, 1, c, -1, plan, lane);
                const uint4 lb_first = lookback_prefetch(lb_addr);   // in flight during the local scan
                const float A1 = sc[0];
                float wdt[RANK > 0 ? RANK : 1];
                if constexpr (RANK > 0) {
#pragma unroll
                    for (int r = 0; r < RANK; ++r) wdt[r] = sc[N + 2 + r];
                }
                float cumA[ITEMS], hloc[ITEMS];
                float P = 1.f, Vv = 0.f;
                auto local_scan = [&](auto tag) {
                    constexpr bool PART = decltype(tag)::value;
#pragma unroll
                    for (int v = 0; v < ITEMS / V; ++v) {
                        float uv[V], dl[V], Bv[V];
                        lds_items<T, V>(su + e0 + v * V, uv);
                        if constexpr (!FUSED) {
                            lds_items<T, V>(su + CL + e0 + v * V, dl);
                        } else {   // dt_proj on the fly: delta = sum_r W[d][r] * dt_lowrank[r], r ascending (weights: smem broadcast)
#pragma unroll
                            for (int k = 0; k < V; ++k) dl[k] = 0.f;
                            if constexpr (RANK > 0) {
#pragma unroll
                                for (int r = 0; r < RANK; ++r) {
                                    float tr[V];
                                    lds_items<T, V>(reinterpret_cast<const T*>(rows + r * ROW_SLOT) + CL + e0 + v * V, tr);
#pragma unroll
                                    for (int k = 0; k < V; ++k) dl[k] = r == 0 ? wdt[0] * tr[k] : fmaf(wdt[r], tr[k], dl[k]);
                                }
                            } else {
                                for (int r = 0; r < R; ++r) {
                                    float tr[V];
                                    lds_items<T, V>(reinterpret_cast<const T*>(rows + r * ROW_SLOT) + CL + e0 + v * V, tr);
                                    const float wr = sc[N + 2 + r];
#pragma unroll
                                    for (int k = 0; k < V; ++k) dl[k] = r == 0 ? wr * tr[k] : fmaf(wr, tr[k], dl[k]);
                                }
                            }
                        }
                        lds_items<T, V>(sB + e0 + v * V, Bv);
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            const int i = v * V + k;
                            float xd = dl[k] + bias;
                            if (p.softplus) xd = softplus_f(xd);
                            float e = decay_m1<kAcc>(xd * A1);
                            float b = xd * uv[k] * Bv[k];
                            if (PART && e0 + i >= len) {   // identity padding so the carried state stays exact
                                e = 0.f;
                                b = 0.f;
                            }
                            decay_step(e, b, P, Vv);
                            hloc[i] = Vv;
                            cumA[i] = P;
                        }
                    }
                };
                if (len < CL) local_scan(std::true_type{});
                else local_scan(std::false_type{});
                if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 12, tc.c);
                warp_scan_fwd(P, Vv, lane);   // (P, Vv): composition of lanes 0..lane
                float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, Vv, 1);
                if (lane == 0) {
                    Pe = 1.f;
                    Ve = 0.f;
                }
                const float Pa = __shfl_sync(FULL, P, 31), Va = __shfl_sync(FULL, Vv, 31);
                float Pp = 1.f, hp = 0.f;
                if (p.lb_dynamic) {   // A/B: classic look-back (timing-dependent association)
                    if (c > 0) {
                        if (lane == 0 && c + 1 < nt) st_desc(aggrow + c, Pa, Va, desc_tag(ep, 1u));
                            const float2 pre = lookback_dynamic(aggrow, 1, c, nt, -1, lane, p.err, ep);
                        Pp = pre.x;
                        hp = pre.y;
                    }
                    if (lane == 0 && c + 1 < nt) st_desc(aggrow + c, Pp * Pa, fmaf(Pa, hp, Va), desc_tag(ep, 2u));
                } else {
                    if (lane == 0 && plan.publish_agg) st_desc(aggrow + c, Pa, Va, desc_tag(ep, DESC_READY));
                    if (plan.nlanes) {
                        const float2 pre = lookback_finish(lb_addr, lb_first, plan.nlanes, lane, p.err, ep);
                        Pp = pre.x;
                        hp = pre.y;
                    }
                    if (lane == 0 && plan.publish_incl) st_desc(inclrow + c, Pp * Pa, fmaf(Pa, hp, Va), desc_tag(ep, DESC_READY));
                }
                if (p.x) {
                    // carries: state and running decay at the end of every kCarry chunk of this tile
#pragma unroll
                    for (int k = 0; k < CPT; ++k) {
                        if (lane == (k + 1) * 32 / CPT - 1 && k * XC < len) {
                            float2* xr = reinterpret_cast<float2*>(p.x) + row * p.nxchunks + (int64_t)c * CPT + k;
                            *xr = make_float2(Pp * P, fmaf(P, hp, Vv));
                        }
                    }
                }
                const float seed = fmaf(Pe, hp, Ve);
                if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 13, tc.c);
#pragma unroll
                for (int v = 0; v < ITEMS / V; ++v) {
                    float uv[V], Cv[V];
                    lds_items<T, V>(su + e0 + v * V, uv);
                    lds_items<T, V>(sC + e0 + v * V, Cv);
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const int i = v * V + k;
                        const float h = fmaf(cumA[i], seed, hloc[i]);
                        y[i] = fmaf(Cv[k], h, Dv * uv[k]);
                    }
                }
            } else {
                // ---------------- general dstate: aggregates first, look-back, then the seeded pass ----------------
                static_assert(N1 || CPT == 1, "general dstate path emits one carry per tile");
                const bool partial = len < CL;
                float uv[ITEMS], dl[ITEMS], du[ITEMS];
                lds_items<T, ITEMS>(su + e0, uv);
                lds_items<T, ITEMS>(su + CL + e0, dl);
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    float xd = dl[i] + bias;
                    if (p.softplus) xd = softplus_f(xd);
                    dl[i] = xd;
                    du[i] = xd * uv[i];
                    y[i] = Dv * uv[i];
                }
                float aggP = 1.f, aggV = 0.f;   // lane n keeps the tile aggregate of state n
                for (int n = 0; n < N; ++n) {
                    const float An = sc[n];
                    float Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + n * CL + e0, Bv);
                    float P = 1.f, Vv = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        float e = decay_m1<kAcc>(dl[i] * An);
                        float b = du[i] * Bv[i];
                        if (partial && e0 + i >= len) {
                            e = 0.f;
                            b = 0.f;
                        }
                        decay_step(e, b, P, Vv);
                    }
                    warp_scan_fwd(P, Vv, lane);
                    const float Pa = __shfl_sync(FULL, P, 31), Va = __shfl_sync(FULL, Vv, 31);
                    if (lane == n) {
                        aggP = Pa;
                        aggV = Va;
                    }
                }
                uint4* aggrow = p.desc + (row * nt) * N;   // [tile][n]
                uint4* inclrow = p.desc_incl + (row * nt) * N;
                if (lane < N && plan.publish_agg) st_desc(aggrow + (int64_t)c * N + lane, aggP, aggV, desc_tag(ep, DESC_READY));
                float preP = 1.f, preV = 0.f;   // lane n: composition of tiles < c for state n
                if (plan.nlanes) {
                    for (int n0 = 0; n0 < N; n0 += 4) {   // four states' descriptors in flight at a time
                        const uint4* addr[4];
                        uint4 first[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            addr[q] = (n0 + q < N) ? lookback_addr(aggrow + n0 + q, inclrow + n0 + q, N, c, -1, plan, lane) : nullptr;
                            first[q] = lookback_prefetch(addr[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (n0 + q < N) {
                                const float2 pre = lookback_finish(addr[q], first[q], plan.nlanes, lane, p.err, ep);
                                if (lane == n0 + q) {
                                    preP = pre.x;
                                    preV = pre.y;
                                }
                            }
                        }
                    }
                }
                if (lane < N) {
                    const float Pi = preP * aggP, hi = fmaf(aggP, preV, aggV);
                    if (plan.publish_incl) st_desc(inclrow + (int64_t)c * N + lane, Pi, hi, desc_tag(ep, DESC_READY));
                    if (p.x) {
                        float2* xr = reinterpret_cast<float2*>(p.x) + (row * p.nxchunks + c) * N + lane;
                        *xr = make_float2(Pi, hi);
                    }
                }
                for (int n = 0; n < N; ++n) {
                    const float An = sc[n];
                    float Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + n * CL + e0, Bv);
                    float cumA[ITEMS], hloc[ITEMS];
                    float P = 1.f, Vv = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        float e = decay_m1<kAcc>(dl[i] * An);
                        float b = du[i] * Bv[i];
                        if (partial && e0 + i >= len) {
                            e = 0.f;
                            b = 0.f;
                        }
                        decay_step(e, b, P, Vv);
                        hloc[i] = Vv;
                        cumA[i] = P;
                    }
                    warp_scan_fwd(P, Vv, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, Vv, 1);
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

            // ---------------- release the stage, then write y straight from registers ----------------
            // Every lane owns ITEMS consecutive outputs (a whole number of 32-byte sectors), so 128-bit stores from
            // registers are sector-complete; no shared-memory staging, no TMA-store drain before the stage can be reused.
            __syncwarp();   // all lanes of this row are done reading the stage
            if (lane == 0) mbar_arrive(&empty[s]);
            if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 14, tc.c);
            OutT* gout = reinterpret_cast<OutT*>(p.out) + tc.b * p.out_bs + d * p.out_ds + l0;
            constexpr int VO = ElemTraits<OutT>::kPerVec;
            if (len == CL && (reinterpret_cast<uintptr_t>(gout) & 15) == 0) {
#pragma unroll
                for (int v = 0; v < ITEMS / VO; ++v) {
                    uint4 raw;
                    OutT* e = reinterpret_cast<OutT*>(&raw);
#pragma unroll
                    for (int k = 0; k < VO; ++k) e[k] = ElemTraits<OutT>::from_f(y[v * VO + k]);
                    reinterpret_cast<uint4*>(gout + e0)[v] = raw;
                }
            } else {
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const int e = e0 + i;
                    if (e < len) gout[e] = ElemTraits<OutT>::from_f(y[i]);
                }
            }
        } else {
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename T, typename OutT, int ITEMS, bool N1, int RANK = 0>
static int launch_fwd(ScanFwdArgs a, int sm_count, cudaStream_t stream) {
    constexpr int NW = kScanWarps;
    constexpr int CL = 32 * ITEMS;
    if constexpr (N1 && RANK == 0 && sizeof(T) == 4) {   // fused dt_proj: own instantiations, the plain kernel keeps its registers
        if (a.R == 3) return launch_fwd<T, OutT, ITEMS, N1, 3>(a, sm_count, stream);
        if (a.R == 5) return launch_fwd<T, OutT, ITEMS, N1, 5>(a, sm_count, stream);
        if (a.R > 0) return launch_fwd<T, OutT, ITEMS, N1, -1>(a, sm_count, stream);
    }
    if (a.R > 0 && (RANK == 0 || a.R > kMaxDtRank || a.R > NW || !a.dt_w)) return BEM_ERR_UNSUPPORTED;   // N = 1, fp32, rank <= 8
    auto kernel = scan_fwd_kernel<T, OutT, ITEMS, NW, N1, RANK>;
    a.nchunks = (a.L + CL - 1) / CL;
    a.RB = (a.Dg + NW - 1) / NW;
    a.RT = a.batch * a.G * a.RB;
    const int64_t total = (int64_t)a.nchunks * a.RT;
    if (total > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    a.total_tiles = (int)total;
    const int64_t ndesc = (int64_t)a.batch * a.dim * a.nchunks * a.N;
    a.desc_incl = a.desc + ndesc;
    const int hdr_bytes = 128 + ((NW * (a.N + 2 + a.R) * 4 + 127) / 128) * 128;
    const int stage_bytes = hdr_bytes + NW * 2 * CL * (int)sizeof(T) + 2 * a.N * CL * (int)sizeof(T);
    // two resident CTAs per SM when two stages fit in half of the shared memory, else one CTA with a deeper ring
    const int budget2 = (227 * 1024) / 2 - 1024;
    int stages;
    if (N1 && 2 * stage_bytes + 256 <= budget2) {
        stages = min(3, (budget2 - 256) / stage_bytes);
    } else {
        stages = min(4, (227 * 1024 - 256) / stage_bytes);
        if (stages < 2) return BEM_ERR_UNSUPPORTED;
    }
    if (const char* ev = getenv("BEM_SCAN_TRACE")) a.trace = atoi(ev);   // stage timeline of CTA 0 (tools/trace_scan.py)
    if (const char* ev = getenv("BEM_LB_DYNAMIC")) a.lb_dynamic = atoi(ev) ? 1 : 0;   // A/B knob (tools/): classic look-back
    if (const char* ev = getenv("BEM_FWD_STAGES")) {   // tuning knob (tools/), not a product interface
        const int v = atoi(ev);
        if (v >= 2 && v * stage_bytes + 512 <= 227 * 1024) stages = v;
    }
    a.stages = stages;
    const int smem_bytes = stages * stage_bytes + stages * 2 * 8 + 64;
    // attribute + occupancy are queried once per (kernel instantiation, shared-memory size, device): they cost
    // microseconds of host time per call, comparable to the kernel itself on the training shapes
    static int cached_smem[64] = {0}, cached_per_sm[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] != smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (NW + 1) * 32, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        cached_per_sm[dev] = per_sm < 1 ? 1 : per_sm;
        cached_smem[dev] = smem_bytes;
    }
    const int grid = min(a.total_tiles, sm_count * cached_per_sm[dev]);
    launch_pdl(kernel, dim3(grid), dim3((NW + 1) * 32), smem_bytes, stream, a);
    return (int)cudaGetLastError();
}

int scan_fwd_dispatch(const ScanFwdArgs& a, int dtype, int out_dtype, int sm_count, cudaStream_t stream) {
    const bool n1 = a.N == 1;
    // dstate >= 2 with rows enough to fill the machine: one CTA per row (scan_rows.cu). Very long sequences over few rows stay
    // with the look-back kernel below, which splits L across CTAs (HD, KD 384, N 16, L 129600: 1.62 ms against 1.70 ms).
    const bool long_few = a.N <= kMaxDstate && a.L >= 32768 && (int64_t)a.batch * a.dim < 4LL * sm_count && !getenv("BEM_SCAN_ROWS");
    if (a.R == 0 && !long_few && scan_rows_preferred(a.batch, a.dim, a.N, sm_count))
        return scan_rows_fwd_dispatch(a, dtype, out_dtype, sm_count, stream);
    if (a.N > kMaxDstate) return BEM_ERR_UNSUPPORTED;
    if (dtype == BEM_F32) {
        if (n1) {
            // A/B and debugging knobs (tools/) select the classic schedule of this file
            static const bool classic = (getenv("BEM_FWD_CLASSIC") && atoi(getenv("BEM_FWD_CLASSIC"))) ||
                                        (getenv("BEM_SCAN_TRACE") && atoi(getenv("BEM_SCAN_TRACE")) == 1) ||
                                        getenv("BEM_LB_DYNAMIC") || getenv("BEM_FWD_ITEMS") || getenv("BEM_FWD_STAGES");
            if (!classic) {
                ScanFwdArgs b = a;
                if (const char* tv = getenv("BEM_SCAN_TRACE")) b.trace = atoi(tv) == 2;
                const int rc = scan_fwd_deferred_dispatch(b, sm_count, stream);   // deferred-finish schedule (scan_fwd_deferred.cu)
                if (rc != BEM_ERR_UNSUPPORTED) return rc;
            }
            const char* ev = getenv("BEM_FWD_ITEMS");   // tuning knob (tools/), not a product interface
            if (ev && atoi(ev) == 12) return launch_fwd<float, float, 12, true>(a, sm_count, stream);
            return launch_fwd<float, float, kFwdItemsF32N1, true>(a, sm_count, stream);
        }
        return launch_fwd<float, float, kItemsF32, false>(a, sm_count, stream);
    }
    if (dtype == BEM_F16) {
        if (out_dtype == BEM_F32)
            return n1 ? launch_fwd<__half, float, kItems16, true>(a, sm_count, stream) : launch_fwd<__half, float, kItems16, false>(a, sm_count, stream);
        return n1 ? launch_fwd<__half, __half, kItems16, true>(a, sm_count, stream) : launch_fwd<__half, __half, kItems16, false>(a, sm_count, stream);
    }
    if (dtype == BEM_BF16) {
        if (out_dtype == BEM_F32)
            return n1 ? launch_fwd<__nv_bfloat16, float, kItems16, true>(a, sm_count, stream)
                      : launch_fwd<__nv_bfloat16, float, kItems16, false>(a, sm_count, stream);
        return n1 ? launch_fwd<__nv_bfloat16, __nv_bfloat16, kItems16, true>(a, sm_count, stream)
                  : launch_fwd<__nv_bfloat16, __nv_bfloat16, kItems16, false>(a, sm_count, stream);
    }
    return BEM_ERR_BAD_ARG;
}

}  // namespace bem

// not part of the ABI: reads (and clears) the stage timeline of the scan forward kernel (tools/trace_scan.py)
extern "C" int bem_dbg_scan_trace(unsigned int* out, int max_records) {
    const int total = bem::STRACE_ROLES * bem::STRACE_PER;
    if (max_records < total) return -1;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, bem::g_scan_trace, (size_t)total * sizeof(uint4));
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, bem::g_scan_trace);
    cudaMemset(sym, 0, (size_t)total * sizeof(uint4));
    return total;
}
