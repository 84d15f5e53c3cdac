// scan_bwd.cu — selective-scan backward for sm_100a.
//
// Replaces selective_scan_bwd_kernel (kernels/selective_scan/csrc/selective_scan/cusoflex/
// selective_scan_bwd_kernel_oflex.cuh:73-289). Same gradients:
//   g_t = C_t dout_t + a_{t+1} g_{t+1}            (reverse scan, :205-210)
//   du = D dout + g delta B;  ddelta = g u B + g A (h_t - b_t);  dA += g delta (h_t - b_t)
//   dB = g delta u;  dC = dout h;  dD += dout u;  ddelta *= sigmoid(delta_raw) when softplus   (:214-257)
// Organisation: same persistent producer/consumer CTA as the forward kernel. The forward state at a
// chunk start comes from the carries `x` written by the forward pass (as in the reference, :200); the reverse scan
// crosses chunks with a decoupled look-back over SUCCESSOR chunks (tiles are handed out last chunk first).
// dB/dC: the reference issues 2*N*L fp32 atomics per channel row (:224-237). Here (dstate == 1) every warp sums its
// rows' contributions in its own shared-memory rows while the CTA walks all rows of a group split; the warps are then
// added up and each (b, g, chunk) slab is written once (plain store when the split covers the whole group).
#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

template <typename T, typename DT, int ITEMS, int NW, bool N1>
__global__ void __launch_bounds__((NW + 1) * 32, (N1 && sizeof(T) == 4) ? 2 : 1) scan_bwd_kernel(const ScanBwdArgs p) {
    constexpr int CL = 32 * ITEMS;
    constexpr bool kAcc = sizeof(T) == 4;   // fp32 inputs: <= 1 ulp decay factors (scan_common.cuh decay_m1)
    constexpr int ROW_SLOT = 2 * CL * (int)sizeof(T) + CL * (int)sizeof(DT);   // [u | delta | dout]; du, ddelta overlay u, delta
    extern __shared__ __align__(128) unsigned char smem[];

    const int N = N1 ? 1 : p.N;
    const int S = p.stages;
    const int bc_bytes = N * CL * (int)sizeof(T);
    const int stage_bytes = NW * ROW_SLOT + 2 * bc_bytes;
    float* red = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);   // [2][NW][CL] (dstate == 1 only)
    const int red_bytes = N1 ? 2 * NW * CL * (int)sizeof(float) : 0;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes + red_bytes);
    uint64_t* empty = full + S;
    int2* tile_slot = reinterpret_cast<int2*>(empty + S);

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

    const int ST = p.ST;
    const int GRS = p.G * p.RS;

    // super tile t -> (chunk, batch, group, row split); chunks are handed out LAST FIRST so that every tile a
    // reverse look-back waits on holds a smaller ticket
    auto decode = [&](int t, int& c, int& b, int& g, int& rs) {
        const int q = t / ST;
        c = p.nchunks - 1 - q;
        const int r = t - q * ST;
        b = r / GRS;
        const int rem = r - b * GRS;
        g = rem / p.RS;
        rs = rem - g * p.RS;
    };

    if (warp == NW) {
        // ======================================= producer warp =======================================
        uint32_t it = 0;
        while (true) {
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(p.ticket, 1u);
            t = __shfl_sync(FULL, t, 0);
            const bool done = t >= (unsigned)p.total_tiles;
            int c = 0, b = 0, g = 0, rs = 0;
            if (!done) decode((int)t, c, b, g, rs);
            const int nsteps = done ? 1 : p.RBS;
            for (int j = 0; j < nsteps; ++j, ++it) {
                const int s = it % S;
                const uint32_t use = it / S;
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
                if (done) {
                    if (lane == 0) {
                        tile_slot[s] = make_int2(-1, 0);
                        mbar_arrive(&full[s]);
                    }
                    break;
                }
                const int l0 = c * CL;
                const int len = min(CL, p.L - l0);
                const int row0 = rs * p.rows_per_split + j * NW;                 // within the group
                const int row_end = min((rs + 1) * p.rows_per_split, p.Dg);
                const int nrows = max(0, min(NW, row_end - row0));
                unsigned char* st = smem + (size_t)s * stage_bytes;
                // jobs: u rows, delta rows, dout rows, then N B rows and N C rows (re-staged every step: L2 hits)
                const int njobs = 3 * nrows + 2 * N;
                uint32_t my_bytes = 0;
                for (int pass = 0; pass < 2; ++pass) {
                    for (int jj = lane; jj < njobs; jj += 32) {
                        const unsigned char* src;
                        unsigned char* dst;
                        int esz;
                        if (jj < 3 * nrows) {
                            const int which = jj / nrows;
                            const int r = jj - which * nrows;
                            const int64_t d = (int64_t)g * p.Dg + row0 + r;
                            unsigned char* slot = st + r * ROW_SLOT;
                            if (which == 0) {
                                src = reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(p.u) + b * p.u_bs + d * p.u_ds + l0);
                                dst = slot;
                                esz = sizeof(T);
                            } else if (which == 1) {
                                src = reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(p.delta) + b * p.dl_bs + d * p.dl_ds + l0);
                                dst = slot + CL * sizeof(T);
                                esz = sizeof(T);
                            } else {
                                src = reinterpret_cast<const unsigned char*>(reinterpret_cast<const DT*>(p.dout) + b * p.do_bs + d * p.do_ds + l0);
                                dst = slot + 2 * CL * sizeof(T);
                                esz = sizeof(DT);
                            }
                        } else {
                            const int k = jj - 3 * nrows;
                            const int isc = k >= N;
                            const int n = isc ? k - N : k;
                            src = reinterpret_cast<const unsigned char*>(
                                isc ? reinterpret_cast<const T*>(p.Cm) + b * p.C_bs + g * p.C_gs + n * p.C_ns + l0
                                    : reinterpret_cast<const T*>(p.Bm) + b * p.B_bs + g * p.B_gs + n * p.B_ns + l0);
                            dst = st + NW * ROW_SLOT + (isc ? bc_bytes : 0) + n * CL * sizeof(T);
                            esz = sizeof(T);
                        }
                        const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
                        const uint32_t vec_bytes = aligned ? ((uint32_t)(len * esz) & ~15u) : 0u;
                        if (pass == 0) {
                            const uint32_t tot_bytes = (uint32_t)(len * esz);
                            if (esz == 4) {
                                for (uint32_t o = vec_bytes; o < tot_bytes; o += 4)
                                    *reinterpret_cast<uint32_t*>(dst + o) = *reinterpret_cast<const uint32_t*>(src + o);
                            } else {
                                for (uint32_t o = vec_bytes; o < tot_bytes; o += 2)
                                    *reinterpret_cast<uint16_t*>(dst + o) = *reinterpret_cast<const uint16_t*>(src + o);
                            }
                            my_bytes += vec_bytes;
                        } else if (vec_bytes) {
                            bulk_g2s(dst, src, vec_bytes, &full[s]);
                        }
                    }
                    if (pass == 0) {
                        uint32_t tot = my_bytes;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
                        __syncwarp();
                        if (lane == 0) {
                            tile_slot[s] = make_int2((int)t, j);
                            if (tot > 0) mbar_arrive_expect_tx(&full[s], tot);
                            else mbar_arrive(&full[s]);
                        }
                        __syncwarp();
                    }
                }
            }
            if (done) break;
        }
        return;
    }

    // ========================================= consumer warps =========================================
    int pend_stage = -1;
    for (uint32_t it = 0;; ++it) {
        const int s = it % S;
        mbar_wait(&full[s], (it / S) & 1, p.err);
        const int2 slot = tile_slot[s];
        if (slot.x < 0) break;
        int c, b, g, rs;
        decode(slot.x, c, b, g, rs);
        const int j = slot.y;
        const int l0 = c * CL;
        const int len = min(CL, p.L - l0);
        const bool partial = len < CL;
        const int row_in_group = rs * p.rows_per_split + j * NW + warp;
        const bool active = row_in_group < min((rs + 1) * p.rows_per_split, p.Dg);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        const int e0 = lane * ITEMS;
        bool drained = false;
        auto drain_prev = [&]() {
            if (!drained && pend_stage >= 0 && lane == 0) {
                bulk_wait_read<0>();
                mbar_arrive(&empty[pend_stage]);
            }
            drained = true;
        };

        if (active) {
            const int64_t d = (int64_t)g * p.Dg + row_in_group;
            const int64_t row = (int64_t)b * p.dim + d;
            unsigned char* rslot = st + warp * ROW_SLOT;
            const T* su = reinterpret_cast<const T*>(rslot);
            const T* sd = su + CL;
            const DT* sdo = reinterpret_cast<const DT*>(rslot + 2 * CL * sizeof(T));
            const T* sB = reinterpret_cast<const T*>(st + NW * ROW_SLOT);
            const T* sC = reinterpret_cast<const T*>(st + NW * ROW_SLOT + bc_bytes);
            const float bias = p.bias ? p.bias[d] : 0.f;
            const float Dv = p.D ? p.D[d] : 0.f;

            float du[ITEMS], dd[ITEMS];   // general-dstate path only (dstate == 1 writes in place)
            float dD_acc = 0.f, dbias_acc = 0.f;

            if constexpr (N1) {
                const float Av = p.A[d * p.A_ds];
                const float A2 = Av;
                float a[ITEMS], h[ITEMS], gl[ITEMS], rp[ITEMS];
                float Pth, Rth;
                {
                    float uv[ITEMS], dl[ITEMS], Bv[ITEMS];
                    lds_items<T, ITEMS>(su + e0, uv);
                    lds_items<T, ITEMS>(sd + e0, dl);
                    lds_items<T, ITEMS>(sB + e0, Bv);
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        float x = dl[i] + bias;
                        if (p.softplus) x = softplus_f(x);
                        float ei = decay_m1<kAcc>(x * A2);
                        float bi = x * uv[i] * Bv[i];
                        if (partial && e0 + i >= len) {
                            ei = 0.f;
                            bi = 0.f;
                        }
                        a[i] = ei;   // decay minus one
                        decay_step(ei, bi, P, V);
                        h[i] = V;    // local inclusive state
                        rp[i] = P;   // local inclusive decay (temporarily)
                    }
                    Pth = P;
                    warp_scan_fwd(P, V, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                    if (lane == 0) {
                        Pe = 1.f;
                        Ve = 0.f;
                    }
                    const float h_in = (c > 0 && p.x) ? p.x[(row * p.nchunks + (c - 1)) * 2 + 1] : 0.f;
                    const float seed = fmaf(Pe, h_in, Ve);
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) h[i] = fmaf(rp[i], seed, h[i]);   // true forward state h_t
                }
                {
                    float Cv[ITEMS], dy[ITEMS];
                    lds_items<T, ITEMS>(sC + e0, Cv);
                    lds_items<DT, ITEMS>(sdo + e0, dy);
                    float r = 0.f, RP = 1.f;
#pragma unroll
                    for (int i = ITEMS - 1; i >= 0; --i) {
                        const float cd = (partial && e0 + i >= len) ? 0.f : Cv[i] * dy[i];
                        gl[i] = cd + r;   // g_t with zero incoming adjoint
                        rp[i] = RP;       // d g_t / d incoming
                        r = fmaf(a[i], gl[i], gl[i]);
                        RP = fmaf(a[i], RP, RP);
                    }
                    Rth = r;
                }
                float P = Pth, R = Rth;
                warp_scan_rev(P, R, lane);
                float Ps = __shfl_down_sync(FULL, P, 1), Rs = __shfl_down_sync(FULL, R, 1);
                if (lane == 31) {
                    Ps = 1.f;
                    Rs = 0.f;
                }
                const float Pa = __shfl_sync(FULL, P, 0), Ra = __shfl_sync(FULL, R, 0);
                uint4* aggrow = p.desc + row * p.nchunks;
                uint4* inclrow = p.desc_incl + row * p.nchunks;
                const LookbackPlan plan = lookback_plan(p.nchunks - 1 - c, p.nchunks);
                if (lane == 0 && plan.publish_agg) st_desc(aggrow + c, Pa, Ra, DESC_READY);
                drain_prev();
                float r_in = 0.f, Psuf = 1.f;
                if (plan.nlanes) {
                    const uint4* lb_addr = lookback_addr(aggrow, inclrow, 1, c, +1, plan, lane);
                    const float2 suf = lookback_finish(lb_addr, lookback_prefetch(lb_addr), plan.nlanes, lane, p.err);
                    Psuf = suf.x;
                    r_in = suf.y;
                }
                if (lane == 0 && plan.publish_incl) st_desc(inclrow + c, Pa * Psuf, fmaf(Pa, r_in, Ra), DESC_READY);
                const float rin_t = fmaf(Ps, r_in, Rs);   // adjoint entering this lane's last position
                {
                    // outputs, one 128-bit vector of T at a time. du / ddelta overwrite u / delta IN PLACE (same lane,
                    // same addresses), dB / dC contributions are accumulated into this warp's rows of `red`.
                    constexpr int VT = ElemTraits<T>::kPerVec;
                    T* s_du = reinterpret_cast<T*>(rslot);
                    T* s_dd = s_du + CL;
                    float* accB = red + warp * CL + e0;
                    float* accC = red + NW * CL + warp * CL + e0;
                    float dA_acc = 0.f;
#pragma unroll
                    for (int v = 0; v < ITEMS / VT; ++v) {
                        float uv[VT], dl[VT], Bv[VT], dy[VT], cB[VT], cC[VT], duv[VT], ddv[VT];
                        lds_items<T, VT>(su + e0 + v * VT, uv);
                        lds_items<T, VT>(sd + e0 + v * VT, dl);
                        lds_items<T, VT>(sB + e0 + v * VT, Bv);
                        lds_items<DT, VT>(sdo + e0 + v * VT, dy);
                        if (j > 0) {
                            lds_items<float, VT>(accB + v * VT, cB);
                            lds_items<float, VT>(accC + v * VT, cC);
                        } else {
#pragma unroll
                            for (int k = 0; k < VT; ++k) cB[k] = cC[k] = 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < VT; ++k) {
                            const int i = v * VT + k;
                            const bool valid = !(partial && e0 + i >= len);
                            const float raw = valid ? dl[k] + bias : 0.f;
                            const float x = p.softplus ? softplus_f(raw) : raw;
                            const float ui = valid ? uv[k] : 0.f;
                            const float dyi = valid ? dy[k] : 0.f;
                            const float Bi = valid ? Bv[k] : 0.f;
                            const float gt = fmaf(rp[i], rin_t, gl[i]);
                            const float bi = x * ui * Bi;
                            const float tt = valid ? gt * (h[i] - bi) : 0.f;   // g_t * a_t * h_{t-1}
                            duv[k] = fmaf(gt * x, Bi, Dv * dyi);
                            float ddl = fmaf(gt * ui, Bi, Av * tt);
                            dA_acc = fmaf(x, tt, dA_acc);
                            cB[k] = fmaf(gt * x, ui, cB[k]);
                            cC[k] = fmaf(dyi, h[i], cC[k]);
                            dD_acc = fmaf(dyi, ui, dD_acc);
                            if (p.softplus) ddl *= softplus_grad_f(raw);
                            if (!valid) ddl = 0.f;
                            dbias_acc += ddl;
                            ddv[k] = ddl;
                        }
                        sts_items<T, VT>(s_du + e0 + v * VT, duv);
                        sts_items<T, VT>(s_dd + e0 + v * VT, ddv);
                        sts_items<float, VT>(accB + v * VT, cB);
                        sts_items<float, VT>(accC + v * VT, cC);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) dA_acc += __shfl_xor_sync(FULL, dA_acc, o);
                    if (lane == 0) atomicAdd(p.dA + d, dA_acc);
                }
            } else {
                // ------------------------------ general dstate ------------------------------
                float uv[ITEMS], dl[ITEMS], dy[ITEMS];
                lds_items<T, ITEMS>(su + e0, uv);
                lds_items<T, ITEMS>(sd + e0, dl);
                lds_items<DT, ITEMS>(sdo + e0, dy);
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const bool valid = !(partial && e0 + i >= len);
                    float x = valid ? dl[i] + bias : 0.f;   // stale shared memory beyond the sequence end may hold NaN patterns
                    if (p.softplus) x = softplus_f(x);
                    dl[i] = x;
                    if (!valid) {
                        uv[i] = 0.f;
                        dy[i] = 0.f;
                    }
                    du[i] = Dv * dy[i];
                    dd[i] = 0.f;
                    dD_acc = fmaf(dy[i], uv[i], dD_acc);
                }
                // pass 1: chunk aggregates of the reverse scan, one state per lane
                float aggP = 1.f, aggR = 0.f;
                for (int n = 0; n < N; ++n) {
                    const float A2 = p.A[d * p.A_ds + n * p.A_ns];
                    float Cv[ITEMS];
                    lds_items<T, ITEMS>(sC + n * CL + e0, Cv);
                    float r = 0.f, RP = 1.f;
#pragma unroll
                    for (int i = ITEMS - 1; i >= 0; --i) {
                        const bool valid = !(partial && e0 + i >= len);
                        const float ei = valid ? decay_m1<kAcc>(dl[i] * A2) : 0.f;
                        const float cd = valid ? Cv[i] * dy[i] : 0.f;
                        const float gsum = cd + r;
                        r = fmaf(ei, gsum, gsum);
                        RP = fmaf(ei, RP, RP);
                    }
                    warp_scan_rev(RP, r, lane);
                    const float Pa = __shfl_sync(FULL, RP, 0), Ra = __shfl_sync(FULL, r, 0);
                    if (lane == n) {
                        aggP = Pa;
                        aggR = Ra;
                    }
                }
                uint4* aggrow = p.desc + (row * p.nchunks) * N;
                uint4* inclrow = p.desc_incl + (row * p.nchunks) * N;
                const LookbackPlan plan = lookback_plan(p.nchunks - 1 - c, p.nchunks);
                if (lane < N && plan.publish_agg) st_desc(aggrow + (int64_t)c * N + lane, aggP, aggR, DESC_READY);
                drain_prev();
                float sufP = 1.f, sufR = 0.f;
                if (plan.nlanes) {
                    for (int n0 = 0; n0 < N; n0 += 4) {   // four states' descriptors in flight at a time
                        const uint4* addr[4];
                        uint4 first[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            addr[q] = (n0 + q < N) ? lookback_addr(aggrow + n0 + q, inclrow + n0 + q, N, c, +1, plan, lane) : nullptr;
                            first[q] = lookback_prefetch(addr[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (n0 + q < N) {
                                const float2 suf = lookback_finish(addr[q], first[q], plan.nlanes, lane, p.err);
                                if (lane == n0 + q) {
                                    sufP = suf.x;
                                    sufR = suf.y;
                                }
                            }
                        }
                    }
                }
                if (lane < N && plan.publish_incl) st_desc(inclrow + (int64_t)c * N + lane, aggP * sufP, fmaf(aggP, sufR, aggR), DESC_READY);
                // pass 2: per state, forward states from the carry, reverse adjoints from the look-back
                for (int n = 0; n < N; ++n) {
                    const float Av = p.A[d * p.A_ds + n * p.A_ns];
                    const float A2 = Av;
                    float a[ITEMS], h[ITEMS], gl[ITEMS], rp[ITEMS], Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + n * CL + e0, Bv);
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const bool valid = !(partial && e0 + i >= len);
                        const float ei = valid ? decay_m1<kAcc>(dl[i] * A2) : 0.f;
                        const float bi = valid ? dl[i] * uv[i] * Bv[i] : 0.f;
                        a[i] = ei;   // decay minus one
                        decay_step(ei, bi, P, V);
                        h[i] = V;
                        rp[i] = P;
                    }
                    float Pth = P;
                    warp_scan_fwd(P, V, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                    if (lane == 0) {
                        Pe = 1.f;
                        Ve = 0.f;
                    }
                    const float h_in = (c > 0 && p.x) ? p.x[((row * p.nchunks + (c - 1)) * N + n) * 2 + 1] : 0.f;
                    const float seed = fmaf(Pe, h_in, Ve);
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) h[i] = fmaf(rp[i], seed, h[i]);
                    float r = 0.f, RP = 1.f;
                    {
                        float Cv[ITEMS];
                        lds_items<T, ITEMS>(sC + n * CL + e0, Cv);
#pragma unroll
                        for (int i = ITEMS - 1; i >= 0; --i) {
                            const bool valid = !(partial && e0 + i >= len);
                            const float cd = valid ? Cv[i] * dy[i] : 0.f;
                            gl[i] = cd + r;
                            rp[i] = RP;
                            r = fmaf(a[i], gl[i], gl[i]);
                            RP = fmaf(a[i], RP, RP);
                        }
                    }
                    float Pr = Pth, Rr = r;
                    warp_scan_rev(Pr, Rr, lane);
                    float Ps = __shfl_down_sync(FULL, Pr, 1), Rs = __shfl_down_sync(FULL, Rr, 1);
                    if (lane == 31) {
                        Ps = 1.f;
                        Rs = 0.f;
                    }
                    const float r_in = __shfl_sync(FULL, sufR, n);
                    const float rin_t = fmaf(Ps, r_in, Rs);
                    float dA_acc = 0.f;
                    float* gB = p.dB + (((int64_t)b * p.G + g) * N + n) * p.L + l0 + e0;
                    float* gC = p.dC + (((int64_t)b * p.G + g) * N + n) * p.L + l0 + e0;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const bool valid = !(partial && e0 + i >= len);
                        const float gt = fmaf(rp[i], rin_t, gl[i]);
                        const float bi = dl[i] * uv[i] * Bv[i];
                        const float tt = valid ? gt * (h[i] - bi) : 0.f;
                        du[i] = fmaf(gt * dl[i], Bv[i], du[i]);
                        dd[i] += fmaf(gt * uv[i], Bv[i], Av * tt);
                        dA_acc = fmaf(dl[i], tt, dA_acc);
                        if (valid) {
                            atomicAdd(gB + i, gt * dl[i] * uv[i]);
                            atomicAdd(gC + i, dy[i] * h[i]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) dA_acc += __shfl_xor_sync(FULL, dA_acc, o);
                    if (lane == 0) atomicAdd(p.dA + d * N + n, dA_acc);
                }
                // softplus chain rule needs the raw delta again
                {
                    float raw[ITEMS];
                    lds_items<T, ITEMS>(sd + e0, raw);
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const bool valid = !(partial && e0 + i >= len);
                        float ddl = dd[i];
                        if (p.softplus) ddl *= softplus_grad_f(raw[i] + bias);
                        if (!valid) ddl = 0.f;
                        dd[i] = ddl;
                        dbias_acc += ddl;
                    }
                }
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

            // ---------------- du / ddelta: in place over the u / delta slots, TMA bulk store ----------------
            drain_prev();
            T* gdu = reinterpret_cast<T*>(p.du) + b * p.du_bs + d * p.du_ds + l0;
            T* gdd = reinterpret_cast<T*>(p.ddelta) + b * p.dd_bs + d * p.dd_ds + l0;
            const bool al_u = (reinterpret_cast<uintptr_t>(gdu) & 15) == 0;
            const bool al_d = (reinterpret_cast<uintptr_t>(gdd) & 15) == 0;
            const uint32_t vb_u = al_u ? ((uint32_t)(len * sizeof(T)) & ~15u) : 0u;
            const uint32_t vb_d = al_d ? ((uint32_t)(len * sizeof(T)) & ~15u) : 0u;
            T* s_du = reinterpret_cast<T*>(rslot);
            T* s_dd = s_du + CL;
            if constexpr (!N1) {
                sts_items<T, ITEMS>(s_du + e0, du);
                sts_items<T, ITEMS>(s_dd + e0, dd);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (vb_u) bulk_s2g(gdu, s_du, vb_u);
                if (vb_d) bulk_s2g(gdd, s_dd, vb_d);
            }
            const int ve_u = vb_u / sizeof(T), ve_d = vb_d / sizeof(T);
            if (ve_u < len || ve_d < len) {   // ragged tail / unaligned rows: this lane's own positions, straight from its slot
                for (int i = 0; i < ITEMS; ++i) {
                    const int e = e0 + i;
                    if (e >= ve_u && e < len) gdu[e] = s_du[e];
                    if (e >= ve_d && e < len) gdd[e] = s_dd[e];
                }
            }
        } else {
            drain_prev();
            if constexpr (N1) {
                if (j == 0) {   // this warp has no row in the first step: its dB/dC rows start at zero
                    float z[ITEMS];
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) z[i] = 0.f;
                    sts_items<float, ITEMS>(red + warp * CL + e0, z);
                    sts_items<float, ITEMS>(red + NW * CL + warp * CL + e0, z);
                }
            }
        }
        if (lane == 0) bulk_commit();
        pend_stage = s;

        if constexpr (N1) {
            if (j == p.RBS - 1) {
                // ---------------- flush dB / dC of this (b, g, split, chunk) slab ----------------
                float* redB = red;
                float* redC = red + NW * CL;
                named_bar_sync(1, NW * 32);
                constexpr int PER_WARP = CL / NW;
                float* gB = p.dB + ((int64_t)b * p.G + g) * p.L + l0;
                float* gC = p.dC + ((int64_t)b * p.G + g) * p.L + l0;
                for (int e = warp * PER_WARP + lane; e < (warp + 1) * PER_WARP; e += 32) {
                    float sb = 0.f, sc = 0.f;
#pragma unroll
                    for (int r = 0; r < NW; ++r) {
                        sb += redB[r * CL + e];
                        sc += redC[r * CL + e];
                    }
                    if (e < len) {
                        if (p.atomic_bc) {
                            atomicAdd(gB + e, sb);
                            atomicAdd(gC + e, sc);
                        } else {
                            gB[e] = sb;
                            gC[e] = sc;
                        }
                    }
                }
                named_bar_sync(1, NW * 32);
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
template <typename T, typename DT, int ITEMS, bool N1>
static int launch_bwd(ScanBwdArgs& a, int sm_count, cudaStream_t stream) {
    constexpr int NW = kScanWarps;
    constexpr int CL = 32 * ITEMS;
    auto kernel = scan_bwd_kernel<T, DT, ITEMS, NW, N1>;
    const int stage_bytes = NW * (2 * CL * (int)sizeof(T) + CL * (int)sizeof(DT)) + 2 * a.N * CL * (int)sizeof(T);
    const int red_bytes = N1 ? 2 * NW * CL * (int)sizeof(float) : 0;
    const int fixed = red_bytes + 512;
    const int budget2 = (227 * 1024) / 2 - 1024;
    const bool two_ok = N1 && sizeof(T) == 4;   // matches __launch_bounds__ of the kernel
    int stages, ctas_per_sm;
    if (two_ok && 2 * stage_bytes + fixed <= budget2) {
        stages = min(4, (budget2 - fixed) / stage_bytes);
        ctas_per_sm = 2;
    } else {
        stages = min(4, (227 * 1024 - fixed) / stage_bytes);
        ctas_per_sm = 1;
        if (stages < 2) return BEM_ERR_UNSUPPORTED;
    }
    a.stages = stages;
    const int smem_bytes = stages * stage_bytes + red_bytes + stages * (2 * 8 + 8) + 64;
    static int cached_smem[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] != smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        cached_smem[dev] = smem_bytes;
    }
    const int grid = min(a.total_tiles, sm_count * ctas_per_sm);
    kernel<<<grid, (NW + 1) * 32, smem_bytes, stream>>>(a);
    return (int)cudaGetLastError();
}

template <typename T, typename DT, int ITEMS>
static int launch_bwd_n(ScanBwdArgs& a, int sm_count, cudaStream_t stream) {
    if (a.N == 1) return launch_bwd<T, DT, ITEMS, true>(a, sm_count, stream);
    return launch_bwd<T, DT, ITEMS, false>(a, sm_count, stream);
}

int scan_bwd_dispatch(ScanBwdArgs& a, int dtype, int dout_dtype, int sm_count, cudaStream_t stream) {
    if (dtype == BEM_F32) return launch_bwd_n<float, float, kItemsF32>(a, sm_count, stream);
    if (dtype == BEM_F16) {
        if (dout_dtype == BEM_F32) return launch_bwd_n<__half, float, kItems16>(a, sm_count, stream);
        return launch_bwd_n<__half, __half, kItems16>(a, sm_count, stream);
    }
    if (dtype == BEM_BF16) {
        if (dout_dtype == BEM_F32) return launch_bwd_n<__nv_bfloat16, float, kItems16>(a, sm_count, stream);
        return launch_bwd_n<__nv_bfloat16, __nv_bfloat16, kItems16>(a, sm_count, stream);
    }
    return BEM_ERR_BAD_ARG;
}

}  // namespace bem
