// scan_fwd_deferred.cu — selective scan forward, fp32, dstate 1: the deferred-finish schedule.
//
// Same algorithm, tile decomposition, descriptors and results as scan_fwd.cu (see there and DESIGN.md 3.1); what changes
// is WHEN a warp finishes a tile. The timeline of the classic schedule (profiles/r01_scan_fwd.md) shows a consumer warp
// waiting ~4300 of ~10400 clk per tile in the look-back: the predecessor tiles of its row run concurrently in other CTAs,
// or sit queued behind another CTA's current tile, so their aggregates are not published yet. Here a warp
//   1. runs pass 1 of tile k+1 — the local scan, which ends by PUBLISHING the tile's aggregate — and only then
//   2. finishes tile k: look-back (its descriptors were requested before step 1 and have had a whole pass to arrive),
//      carries, pass 2, stores.
// Every aggregate is therefore published as early as the data allows, independent of any look-back, and the look-back
// of tile k overlaps useful work. Pass 1 leaves per position  alpha_i = C_i * P_i  and  beta_i = C_i * h_i + D * u_i
// (P_i / h_i: running decay / state from the lane's first position) in place of u_i / delta_i in the stage, so pass 2 is
// y_i = alpha_i * seed + beta_i and nothing but six scalars per lane is carried in registers between the two steps.
// A tile lives in its stage for two steps, hence three or four stages per CTA: 384-position tiles (12 per lane — whose
// 48-byte lane stride is also free of LDS bank conflicts), 2 CTAs per SM. Two producer warps split the bulk copies of a
// tile (the issue of ~18 copies at ~140 clk each by one warp was the next limiter).
// Deadlock freedom: a look-back only ever waits for the pass 1 of tiles with smaller tickets, and a warp runs pass 1 of
// its newest tile before any look-back; the tile with the smallest outstanding ticket waits for nothing but its data.
#include <cstdlib>
#include <type_traits>

#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

namespace {
constexpr int D_ITEMS = 12, D_NW = 8, D_CL = 32 * D_ITEMS, D_V = 4;
constexpr int D_ROW_SLOT = 2 * D_CL * 4;   // [u chunk | delta chunk] of one channel row, later [alpha | beta]
constexpr int D_THREADS = (D_NW + 2) * 32;
static_assert(D_CL == kCarryF32, "one carry of `x` per tile");

// stage timeline of CTA 0 (tools/trace_scan.py, env BEM_SCAN_TRACE=2): one region of records per traced warp, plain stores
constexpr int DT_ROLES = 4, DT_PER = 2048;
__device__ uint4 g_trace_d[DT_ROLES * DT_PER];
struct Tracer {
    uint32_t n = 0;
    __device__ __forceinline__ void operator()(int on, int role, uint32_t tag, uint32_t arg) {
        if (on && blockIdx.x == 0 && n < DT_PER) g_trace_d[role * DT_PER + n++] = make_uint4(tag, arg, (uint32_t)clock64(), 1u);
    }
};

struct Pending {   // what a lane keeps of a tile between pass 1 and its finish
    int s, c, len, active;
    int nlanes, publish_incl;          // look-back plan of the tile (lookback_plan)
    const uint4* lb_addr;              // this lane's look-back descriptor (nullptr: idle lane)
    uint4* incl;                       // the tile's inclusive descriptor
    float2* carry;                     // the tile's carry in `x` (nullptr: not requested)
    float* gout;                       // the row's output at the tile start
    float Pe, Ve, Pa, Va;
};
}  // namespace

// RANK: fused dt_proj rank. 0 = delta per channel row; > 0 compile-time rank; -1 = rank from the arguments (<= kMaxDtRank).
// SP: delta_softplus as a compile-time constant (1 / 0) for the plain kernel, -1 = read from the arguments: a (uniform) branch
// around softplus_f in every element fences the unrolled elements off from each other in the instruction schedule.
template <int RANK, int SP = -1, bool TRACE = false>
__global__ void __launch_bounds__(D_THREADS, 2) scan_fwd_deferred_kernel(const ScanFwdArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NW = D_NW, CL = D_CL, ITEMS = D_ITEMS, V = D_V, ROW_SLOT = D_ROW_SLOT;
    constexpr bool FUSED = RANK != 0;
    extern __shared__ __align__(128) unsigned char smem[];
    const int R = RANK > 0 ? RANK : (RANK < 0 ? p.R : 0);
    const int NSC = 3 + R;                                                // scalars per row: A, D, bias, W_dt[R]
    const int S = p.stages;
    const int hdr_bytes = 128 + ((NW * NSC * 4 + 127) / 128) * 128;      // TileCoord | per-row scalars
    const int bc_bytes = CL * 4;
    const int stage_bytes = hdr_bytes + NW * ROW_SLOT + 2 * bc_bytes + R * CL * 4;   // ... | B | C | low-rank dt rows
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;
    uint64_t* hdr_ready = empty + S;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Tracer tr;
    constexpr int trace_on = TRACE ? 1 : 0;   // the timeline instantiation is separate: no trace predicates in the product kernel
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 2);        // one expect_tx arrival per producer warp
            mbar_init(&empty[s], NW);
            mbar_init(&hdr_ready[s], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp == NW) {
        // ============================ producer 0: tickets, tile header, scalars, u rows ============================
        const int RT = p.RT, GRB = p.G * p.RB;
        constexpr int kMaxSc = (NW * (3 + kMaxDtRank) + 31) / 32;
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
                    mbar_arrive(&hdr_ready[s]);
                    mbar_arrive(&full[s]);
                    if (t == (unsigned)p.total_tiles + gridDim.x - 1) {   // last failing draw of the launch: re-arm the workspace
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
            float scv[kMaxSc];
#pragma unroll
            for (int q = 0; q < kMaxSc; ++q) {
                const int i = lane + 32 * q;
                float v = 0.f;
                if (i < tc.nrows * NSC) {
                    const int rr = i / NSC, k = i - rr * NSC;
                    const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + rr;
                    if (k == 0) v = p.A[d * p.A_ds];
                    else if (k == 1) v = p.D ? p.D[d] : 0.f;
                    else if (k == 2) v = p.bias ? p.bias[d] : 0.f;
                    else v = p.dt_w[d * R + (k - 3)];
                }
                scv[q] = v;
            }
            // this lane's copy job: u row `lane` of the tile
            const float* src = nullptr;
            float* dst = nullptr;
            uint32_t vec_bytes = 0;
            if (lane < tc.nrows) {
                const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + lane;
                src = reinterpret_cast<const float*>(p.u) + tc.b * p.u_bs + d * p.u_ds + l0;
                vec_bytes = (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? ((uint32_t)(tc.len * 4) & ~15u) : 0u;
            }
            if (lane == 0) tr(trace_on, 0, 1, t);
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
            if (lane == 0) tr(trace_on, 0, 2, t);
            unsigned int t_next = 0;
            if (lane == 0) t_next = atomicAdd(p.ticket, 1u);   // drawn only once the slot is free (see scan_fwd.cu)
            if (lane == 0) *hdr = tc;
            float* sc = reinterpret_cast<float*>(st + 128);
#pragma unroll
            for (int q = 0; q < kMaxSc; ++q) {
                const int i = lane + 32 * q;
                if (i < tc.nrows * NSC) sc[i] = scv[q];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr_ready[s]);          // producer 1 may read the header now
            unsigned char* rows = st + hdr_bytes;
            if (lane < tc.nrows) {
                dst = reinterpret_cast<float*>(rows + lane * ROW_SLOT);
                for (int e = vec_bytes / 4; e < tc.len; ++e) dst[e] = src[e];   // ragged tail / unaligned row: plain loads
            }
            uint32_t tot = vec_bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            __syncwarp();
            if (lane == 0) {
                if (tot > 0) mbar_arrive_expect_tx(&full[s], tot);
                else mbar_arrive(&full[s]);
            }
            __syncwarp();
            if (vec_bytes) bulk_g2s(dst, src, vec_bytes, &full[s]);
            if (lane == 0) tr(trace_on, 0, 3, t);
            t = __shfl_sync(FULL, t_next, 0);
            if (++s == S) {
                s = 0;
                ++use;
            }
        }
        return;
    }
    if (warp == NW + 1) {
        // ============================ producer 1: delta (or low-rank dt) rows, B, C ============================
        int s = 0;
        uint32_t use = 0;
        while (true) {
            unsigned char* st = smem + (size_t)s * stage_bytes;
            if (lane == 0) tr(trace_on, 3, 4, use);
            mbar_wait(&hdr_ready[s], use & 1, p.err);
            const TileCoord tc = *reinterpret_cast<const TileCoord*>(st);
            if (tc.nrows < 0) {
                if (lane == 0) mbar_arrive(&full[s]);
                break;
            }
            if (lane == 0) tr(trace_on, 3, 5, tc.c);
            const int l0 = tc.c * CL;
            unsigned char* rows = st + hdr_bytes;
            const int nd = FUSED ? R : tc.nrows;
            const float* src = nullptr;
            float* dst = nullptr;
            uint32_t vec_bytes = 0;
            if (lane < nd + 2) {
                if (lane < nd) {
                    if (FUSED) {
                        src = reinterpret_cast<const float*>(p.delta) + tc.b * p.dl_bs + tc.g * p.dl_gs + lane * p.dl_ds + l0;
                        dst = reinterpret_cast<float*>(rows + NW * ROW_SLOT + 2 * bc_bytes) + lane * CL;
                    } else {
                        const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + lane;
                        src = reinterpret_cast<const float*>(p.delta) + tc.b * p.dl_bs + d * p.dl_ds + l0;
                        dst = reinterpret_cast<float*>(rows + lane * ROW_SLOT) + CL;
                    }
                } else if (lane == nd) {
                    src = reinterpret_cast<const float*>(p.Bm) + tc.b * p.B_bs + tc.g * p.B_gs + l0;
                    dst = reinterpret_cast<float*>(rows + NW * ROW_SLOT);
                } else {
                    src = reinterpret_cast<const float*>(p.Cm) + tc.b * p.C_bs + tc.g * p.C_gs + l0;
                    dst = reinterpret_cast<float*>(rows + NW * ROW_SLOT + bc_bytes);
                }
                vec_bytes = (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? ((uint32_t)(tc.len * 4) & ~15u) : 0u;
                for (int e = vec_bytes / 4; e < tc.len; ++e) dst[e] = src[e];
            }
            uint32_t tot = vec_bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            __syncwarp();
            if (lane == 0) {
                if (tot > 0) mbar_arrive_expect_tx(&full[s], tot);
                else mbar_arrive(&full[s]);
            }
            __syncwarp();
            if (vec_bytes) bulk_g2s(dst, src, vec_bytes, &full[s]);
            if (lane == 0) tr(trace_on, 3, 6, tc.c);
            if (++s == S) {
                s = 0;
                ++use;
            }
        }
        return;
    }

    // ============================================ consumer warps ============================================
    const int nt = p.nchunks;
    const int e0 = lane * ITEMS;
    Pending prev;
    prev.s = -1;
    prev.active = 0;
    uint32_t ep = 0;

    // finish of a tile whose pass 1 is done: look-back (descriptor `first` was requested a whole pass ago), publish the
    // inclusive value / carry, pass 2 from the alpha / beta left in the stage, release the stage, store y
    auto finish = [&](const Pending& q, uint4 lb_first) {
        unsigned char* st = smem + (size_t)q.s * stage_bytes;
        if (!q.active) {
            if (lane == 0) mbar_arrive(&empty[q.s]);
            return;
        }
        float Pp = 1.f, hp = 0.f;
        if (q.nlanes) {
            const float2 pre = lookback_finish(q.lb_addr, lb_first, q.nlanes, lane, p.err, ep);
            Pp = pre.x;
            hp = pre.y;
        }
        if (lane == 0 && q.publish_incl) st_desc(q.incl, Pp * q.Pa, fmaf(q.Pa, hp, q.Va), desc_tag(ep, DESC_READY));
        if (lane == 31 && q.carry) *q.carry = make_float2(Pp * q.Pa, fmaf(q.Pa, hp, q.Va));
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 13, q.c);
        const float seed = fmaf(q.Pe, hp, q.Ve);
        const float* sa = reinterpret_cast<const float*>(st + hdr_bytes + warp * ROW_SLOT) + e0;
        float al[ITEMS], be[ITEMS], y[ITEMS];
        lds_items<float, ITEMS>(sa, al);
        lds_items<float, ITEMS>(sa + CL, be);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) y[i] = fmaf(al[i], seed, be[i]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[q.s]);
        float* gout = q.gout;
        if (q.len == CL && (reinterpret_cast<uintptr_t>(gout) & 15) == 0) {
#pragma unroll
            for (int v = 0; v < ITEMS / V; ++v)
                reinterpret_cast<float4*>(gout + e0)[v] = make_float4(y[v * 4], y[v * 4 + 1], y[v * 4 + 2], y[v * 4 + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                if (e0 + i < q.len) gout[e0 + i] = y[i];
        }
    };

    int s = -1;
    uint32_t phase = 1;
    while (true) {
        if (++s == S) s = 0;
        if (s == 0) phase ^= 1;
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 10, 0);
        mbar_wait(&full[s], phase, p.err);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        const TileCoord tc = *reinterpret_cast<const TileCoord*>(st);
        ep = tc.nrows < 0 ? ep : tc.epoch;
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 11, tc.c);
        // request the previous tile's look-back descriptors now: they are in flight during this tile's pass 1
        uint4 lb_first = make_uint4(0u, 0u, 0u, 0u);
        if (prev.s >= 0 && prev.active) lb_first = lookback_prefetch(prev.lb_addr);
        if (tc.nrows < 0) {
            if (prev.s >= 0) finish(prev, lb_first);
            break;
        }
        Pending cur;
        cur.s = s;
        cur.c = tc.c;
        cur.len = tc.len;
        cur.active = warp < tc.nrows;
        cur.Pe = 1.f;
        cur.Ve = 0.f;
        cur.Pa = 1.f;
        cur.Va = 0.f;
        cur.nlanes = cur.publish_incl = 0;
        cur.lb_addr = nullptr;
        cur.incl = nullptr;
        cur.carry = nullptr;
        cur.gout = nullptr;
        if (cur.active) {
            // ------------------------------ pass 1: local scan, alpha / beta in place ------------------------------
            const int64_t d = (int64_t)tc.g * p.Dg + tc.row0 + warp;
            const int row = tc.b * p.dim + (int)d;                      // (batch, dim) row index: fits 32 bits (host check)
            const LookbackPlan plan = lookback_plan(tc.c, nt);
            uint4* aggrow = p.desc + (uint32_t)(row * nt);             // 32-bit index arithmetic, one widening add per pointer
            uint4* inclrow = p.desc_incl + (uint32_t)(row * nt);
            cur.nlanes = plan.nlanes;
            cur.publish_incl = plan.publish_incl;
            cur.lb_addr = lookback_addr(aggrow, inclrow, 1, tc.c, -1, plan, lane);
            cur.incl = inclrow + tc.c;
            cur.carry = p.x ? reinterpret_cast<float2*>(p.x) + (uint32_t)(row * p.nxchunks + tc.c) : nullptr;
            cur.gout = reinterpret_cast<float*>(p.out) + tc.b * p.out_bs + d * p.out_ds + (int64_t)tc.c * CL;
            const float* sc = reinterpret_cast<const float*>(st + 128) + warp * NSC;
            unsigned char* rows = st + hdr_bytes;
            float* su = reinterpret_cast<float*>(rows + warp * ROW_SLOT) + e0;
            const float* sB = reinterpret_cast<const float*>(rows + NW * ROW_SLOT) + e0;
            const float* sC = sB + CL;
            const float* slr = reinterpret_cast<const float*>(rows + NW * ROW_SLOT + 2 * bc_bytes) + e0;
            const float A1 = sc[0], Dv = sc[1], bias = sc[2];
            float wdt[RANK > 0 ? RANK : 1];
            if constexpr (RANK > 0) {
#pragma unroll
                for (int r = 0; r < RANK; ++r) wdt[r] = sc[3 + r];
            }
            float P = 1.f, Vv = 0.f;
            auto local_scan = [&](auto tag) {
            constexpr bool PART = decltype(tag)::value;   // ragged last tile of a row: identity padding past `len`
#pragma unroll
            for (int v = 0; v < ITEMS / V; ++v) {
                float uv[V], dl[V], Bv[V], Cv[V], al[V], be[V];
                lds_items<float, V>(su + v * V, uv);
                if constexpr (!FUSED) {
                    lds_items<float, V>(su + CL + v * V, dl);
                } else if constexpr (RANK > 0) {
#pragma unroll
                    for (int r = 0; r < RANK; ++r) {
                        float tr[V];
                        lds_items<float, V>(slr + r * CL + v * V, tr);
#pragma unroll
                        for (int k = 0; k < V; ++k) dl[k] = r == 0 ? wdt[0] * tr[k] : fmaf(wdt[r], tr[k], dl[k]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < V; ++k) dl[k] = 0.f;
                    for (int r = 0; r < R; ++r) {
                        float tr[V];
                        lds_items<float, V>(slr + r * CL + v * V, tr);
                        const float wr = sc[3 + r];
#pragma unroll
                        for (int k = 0; k < V; ++k) dl[k] = r == 0 ? wr * tr[k] : fmaf(wr, tr[k], dl[k]);
                    }
                }
                lds_items<float, V>(sB + v * V, Bv);
                lds_items<float, V>(sC + v * V, Cv);
                // the element math that does not depend on the recurrence (softplus, decay, b, D u) runs two positions at a
                // time on the packed fp32 pipe (scan_common.cuh); the recurrence and alpha / beta stay scalar
#pragma unroll
                for (int k = 0; k < V; k += 2) {
                    f32x2 xd = add2(pk2(dl[k], dl[k + 1]), splat2(bias));
                    const bool sp_on = SP >= 0 ? SP == 1 : p.softplus != 0;
                    if (sp_on) {
                        float x0, x1;
                        upk2(xd, x0, x1);
                        xd = softplus2(x0, x1);
                    }
                    f32x2 e2 = decay_m1_2(mul2(xd, splat2(A1)));
                    f32x2 b2 = mul2(mul2(xd, pk2(uv[k], uv[k + 1])), pk2(Bv[k], Bv[k + 1]));
                    const f32x2 du2 = mul2(splat2(Dv), pk2(uv[k], uv[k + 1]));
                    float ee[2], bb[2], du[2];
                    upk2(e2, ee[0], ee[1]);
                    upk2(b2, bb[0], bb[1]);
                    upk2(du2, du[0], du[1]);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (PART && e0 + v * V + k + j >= tc.len) {   // identity padding so the carried state stays exact
                            ee[j] = 0.f;
                            bb[j] = 0.f;
                        }
                        decay_step(ee[j], bb[j], P, Vv);
                        al[k + j] = Cv[k + j] * P;
                        be[k + j] = fmaf(Cv[k + j], Vv, du[j]);
                    }
                }
                sts_items<float, V>(su + v * V, al);
                sts_items<float, V>(su + CL + v * V, be);
            }
            };
            if (tc.len < CL) local_scan(std::true_type{});
            else local_scan(std::false_type{});
            warp_scan_fwd(P, Vv, lane);   // (P, Vv): composition of lanes 0..lane
            cur.Pe = __shfl_up_sync(FULL, P, 1);
            cur.Ve = __shfl_up_sync(FULL, Vv, 1);
            if (lane == 0) {
                cur.Pe = 1.f;
                cur.Ve = 0.f;
            }
            cur.Pa = __shfl_sync(FULL, P, 31);
            cur.Va = __shfl_sync(FULL, Vv, 31);
            if (lane == 0 && plan.publish_agg) st_desc(aggrow + tc.c, cur.Pa, cur.Va, desc_tag(ep, DESC_READY));
        }
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 12, tc.c);
        if (prev.s >= 0) finish(prev, lb_first);
        if (lane == 0 && (warp == 0 || warp == 5)) tr(trace_on, warp == 0 ? 1 : 2, 14, tc.c);
        prev = cur;
    }
}

template <int RANK>
static int launch_deferred(ScanFwdArgs a, int sm_count, cudaStream_t stream) {
    constexpr int NW = D_NW, CL = D_CL;
    a.nchunks = (a.L + CL - 1) / CL;
    a.RB = (a.Dg + NW - 1) / NW;
    a.RT = a.batch * a.G * a.RB;
    const int64_t total = (int64_t)a.nchunks * a.RT;
    if (total > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    a.total_tiles = (int)total;
    a.desc_incl = a.desc + (int64_t)a.batch * a.dim * a.nchunks;
    if ((int64_t)a.batch * a.dim * a.nchunks >= (1ll << 31)) return BEM_ERR_UNSUPPORTED;   // 32-bit descriptor indices in the kernel
    if (a.R > 0 && (RANK == 0 || a.R > kMaxDtRank || !a.dt_w)) return BEM_ERR_UNSUPPORTED;
    const int hdr_bytes = 128 + ((NW * (3 + a.R) * 4 + 127) / 128) * 128;
    const int stage_bytes = hdr_bytes + NW * D_ROW_SLOT + 2 * CL * 4 + a.R * CL * 4;
    const int budget = (227 * 1024) / 2 - 1024;          // two CTAs per SM
    int stages = (budget - 128) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 3) return BEM_ERR_UNSUPPORTED;
    a.stages = stages;
    const int smem_bytes = stages * stage_bytes + stages * 3 * 8 + 64;
    auto kernel = scan_fwd_deferred_kernel<RANK, -1, false>;
    if (RANK == 0) kernel = a.trace ? scan_fwd_deferred_kernel<0, -1, true>
                                    : (a.softplus ? scan_fwd_deferred_kernel<0, 1, false> : scan_fwd_deferred_kernel<0, 0, false>);
    // the attribute belongs to a kernel FUNCTION: one cache slot per variant this launcher may pick
    const int variant = RANK != 0 ? 0 : (a.trace ? 3 : (a.softplus ? 1 : 2));
    static int cached_smem_v[4][64] = {{0}}, cached_per_sm_v[4][64] = {{0}};
    int* cached_smem = cached_smem_v[variant];
    int* cached_per_sm = cached_per_sm_v[variant];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] != smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, D_THREADS, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        cached_per_sm[dev] = per_sm < 1 ? 1 : per_sm;
        cached_smem[dev] = smem_bytes;
    }
    const int grid = min(a.total_tiles, sm_count * cached_per_sm[dev]);
    launch_pdl(kernel, dim3(grid), dim3(D_THREADS), smem_bytes, stream, a);
    return (int)cudaGetLastError();
}

// fp32, dstate 1 (plain or fused dt_proj). Returns BEM_ERR_UNSUPPORTED when the configuration is not built here, in which
// case the caller falls back to the classic schedule.
int scan_fwd_deferred_dispatch(const ScanFwdArgs& a, int sm_count, cudaStream_t stream) {
    if (a.N != 1) return BEM_ERR_UNSUPPORTED;
    if (a.R == 0) return launch_deferred<0>(a, sm_count, stream);
    if (a.R == 3) return launch_deferred<3>(a, sm_count, stream);
    if (a.R == 5) return launch_deferred<5>(a, sm_count, stream);
    return launch_deferred<-1>(a, sm_count, stream);
}

}  // namespace bem

// not part of the ABI: reads (and clears) the stage timeline of the deferred-finish kernel (tools/trace_scan.py)
extern "C" int bem_dbg_scan_deferred_trace(unsigned int* out, int max_records) {
    const int total = bem::DT_ROLES * bem::DT_PER;
    if (max_records < total) return -1;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, bem::g_trace_d, (size_t)total * sizeof(uint4));
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, bem::g_trace_d);
    cudaMemset(sym, 0, (size_t)total * sizeof(uint4));
    return total;
}
