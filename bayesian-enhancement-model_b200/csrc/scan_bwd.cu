// scan_bwd.cu — selective-scan backward for sm_100a.
//
// Replaces selective_scan_bwd_kernel (kernels/selective_scan/csrc/selective_scan/cusoflex/
// selective_scan_bwd_kernel_oflex.cuh:73-289). Same gradients:
//   g_t = C_t dout_t + a_{t+1} g_{t+1}            (reverse scan, :205-210)
//   du = D dout + g delta B;  ddelta = g u B + g A (h_t - b_t);  dA += g delta (h_t - b_t)
//   dB = g delta u;  dC = dout h;  dD += dout u;  ddelta *= sigmoid(delta_raw) when softplus   (:214-257)
// Organisation: same persistent producer/consumer CTA as the forward kernel. The producer stages u / delta /
// dout rows and the B / C chunk with TMA bulk copies and publishes, per tile, the coordinates, the rows' scalars
// (A, D, bias) and the forward state at the tile start (the carries `x` written by the forward pass, as in the
// reference, :200). The reverse scan crosses tiles with the deterministic decoupled look-back over SUCCESSOR tiles
// (tickets are handed out last chunk first). du / ddelta leave straight from registers (128-bit stores).
// dB/dC: the reference issues 2*N*L fp32 atomics per channel row (:224-237). Here (dstate == 1) every warp sums its
// rows' contributions in its own shared-memory rows while the CTA walks all rows of a group split; the warps are then
// added up and each (b, g, chunk) slab is written once (plain store when the split covers the whole group).
#include <type_traits>

#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

// SP: delta_softplus at compile time (1 / 0; fp32 dstate-1 instantiations) or -1 = read from the arguments (see scan_fwd_deferred.cu)
template <typename T, typename DT, int ITEMS, int NW, bool N1, int SP = -1>
__global__ void __launch_bounds__((NW + 1) * 32, (N1 && sizeof(T) == 4) ? 2 : 1) scan_bwd_kernel(const ScanBwdArgs p) {
    pdl_trigger();
    pdl_wait();
    const bool softplus_on = SP >= 0 ? SP == 1 : p.softplus != 0;
    constexpr int CL = 32 * ITEMS;
    constexpr bool kAcc = sizeof(T) == 4;   // fp32 inputs: full-precision decay rate (scan_common.cuh decay_m1)
    constexpr int ROW_SLOT = 2 * CL * (int)sizeof(T) + CL * (int)sizeof(DT);   // [u | delta | dout]
    constexpr int VT = ElemTraits<T>::kPerVec;
    extern __shared__ __align__(128) unsigned char smem[];

    const int N = N1 ? 1 : p.N;
    const int S = p.stages;
    const int nsc = 2 * N + 2;   // per-row scalars: A[N], D, bias, h_in[N]
    const int bc_bytes = N * CL * (int)sizeof(T);
    const int hdr_bytes = 128 + ((NW * nsc * 4 + 127) / 128) * 128;
    const int stage_bytes = hdr_bytes + NW * ROW_SLOT + 2 * bc_bytes;
    float* red = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);   // [2][NW][CL] (dstate == 1 only)
    const int red_bytes = N1 ? 2 * NW * CL * (int)sizeof(float) : 0;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes + red_bytes);
    uint64_t* empty = full + S;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nt = p.nchunks;

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
        const int ST = p.ST;
        const int GRS = p.G * p.RS;
        constexpr int kMaxSc = (NW * (2 * kMaxDstate + 2) + 31) / 32;
        // launch epoch of the descriptors: read before this CTA's first draw (the last drawer of the launch bumps it)
        const uint32_t epoch = *reinterpret_cast<volatile unsigned int*>(p.ticket + 2);
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(p.ticket, 1u);
        t = __shfl_sync(FULL, t, 0);
        int s = 0;
        uint32_t use = 0;
        while (true) {
            if (t >= (unsigned)p.total_tiles) {
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
                if (lane == 0) {
                    reinterpret_cast<TileCoord*>(smem + (size_t)s * stage_bytes)->nrows = -1;
                    mbar_arrive(&full[s]);
                    // the last failing draw of the launch re-arms the workspace (see scan_fwd.cu)
                    if (t == (unsigned)p.total_tiles + gridDim.x - 1) {
                        p.ticket[2] = (epoch + 1) & 0x3fffffffu;
                        p.ticket[0] = 0u;
                    }
                }
                break;
            }
            // super tile t -> (chunk, batch, group, row split); chunks are handed out LAST FIRST so that every tile a
            // reverse look-back waits on holds a smaller ticket
            const int q = (int)t / ST;
            const int c = nt - 1 - q;
            const int r = (int)t - q * ST;
            const int b = r / GRS;
            const int rem = r - b * GRS;
            const int g = rem / p.RS;
            const int rs = rem - g * p.RS;
            const int l0 = c * CL;
            const int len = min(CL, p.L - l0);
            unsigned int t_next = 0;
            for (int j = 0; j < p.RBS; ++j) {
                unsigned char* st = smem + (size_t)s * stage_bytes;
                TileCoord tc;
                tc.c = c;
                tc.b = b;
                tc.g = g;
                tc.row0 = rs * p.rows_per_split + j * NW;   // within the group
                const int row_end = min((rs + 1) * p.rows_per_split, p.Dg);
                tc.nrows = max(0, min(NW, row_end - tc.row0));
                tc.len = len;
                tc.aux0 = j;
                tc.aux1 = (j == p.RBS - 1) ? 1 : 0;
                tc.epoch = epoch;
                // per-row scalars, requested before blocking on the slot
                float scv[kMaxSc];
#pragma unroll
                for (int qq = 0; qq < kMaxSc; ++qq) {
                    const int i = lane + 32 * qq;
                    float v = 0.f;
                    if (i < tc.nrows * nsc) {
                        const int rr = i / nsc, k = i - rr * nsc;
                        const int64_t d = (int64_t)g * p.Dg + tc.row0 + rr;
                        if (k < N) v = p.A[d * p.A_ds + k * p.A_ns];
                        else if (k == N) v = p.D ? p.D[d] : 0.f;
                        else if (k == N + 1) v = p.bias ? p.bias[d] : 0.f;
                        else if (c > 0 && p.x) v = p.x[((((int64_t)b * p.dim + d) * nt + (c - 1)) * N + (k - N - 2)) * 2 + 1];
                    }
                    scv[qq] = v;
                }
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.err);
                if (j == p.RBS - 1 && lane == 0) t_next = atomicAdd(p.ticket, 1u);
                if (lane == 0) *reinterpret_cast<TileCoord*>(st) = tc;
                float* sc = reinterpret_cast<float*>(st + 128);
#pragma unroll
                for (int qq = 0; qq < kMaxSc; ++qq) {
                    const int i = lane + 32 * qq;
                    if (i < tc.nrows * nsc) sc[i] = scv[qq];
                }
                unsigned char* rows = st + hdr_bytes;
                // jobs: u rows, delta rows, dout rows, then N B rows and N C rows (re-staged every step: L2 hits)
                const int nrows = tc.nrows;
                const int njobs = 3 * nrows + 2 * N;
                uint32_t my_bytes = 0;
                for (int pass = 0; pass < 2; ++pass) {
                    for (int jj = lane; jj < njobs; jj += 32) {
                        const unsigned char* src;
                        unsigned char* dst;
                        int esz;
                        if (jj < 3 * nrows) {
                            const int which = jj / nrows;
                            const int rr = jj - which * nrows;
                            const int64_t d = (int64_t)g * p.Dg + tc.row0 + rr;
                            unsigned char* slot = rows + rr * ROW_SLOT;
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
                            dst = rows + NW * ROW_SLOT + (isc ? bc_bytes : 0) + n * CL * sizeof(T);
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
                            if (tot > 0) mbar_arrive_expect_tx(&full[s], tot);
                            else mbar_arrive(&full[s]);
                        }
                        __syncwarp();
                    }
                }
                if (++s == S) {
                    s = 0;
                    ++use;
                }
            }
            t = __shfl_sync(FULL, t_next, 0);
        }
        return;
    }

    // ========================================= consumer warps =========================================
    int s = -1;
    uint32_t phase = 1;
    while (true) {
        if (++s == S) s = 0;
        if (s == 0) phase ^= 1;
        mbar_wait(&full[s], phase, p.err);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        const TileCoord tc = *reinterpret_cast<const TileCoord*>(st);
        const uint32_t ep = tc.epoch;
        if (tc.nrows < 0) break;
        const int c = tc.c, b = tc.b, g = tc.g;
        const int j = tc.aux0;
        const int l0 = c * CL;
        const int len = tc.len;
        const bool partial = len < CL;
        const bool active = warp < tc.nrows;
        const int e0 = lane * ITEMS;

        if (active) {
            const int64_t d = (int64_t)g * p.Dg + tc.row0 + warp;
            const int64_t row = (int64_t)b * p.dim + d;
            const float* sc = reinterpret_cast<const float*>(st + 128) + warp * nsc;
            unsigned char* rows = st + hdr_bytes;
            unsigned char* rslot = rows + warp * ROW_SLOT;
            const T* su = reinterpret_cast<const T*>(rslot);
            const T* sd = su + CL;
            const DT* sdo = reinterpret_cast<const DT*>(rslot + 2 * CL * sizeof(T));
            const T* sB = reinterpret_cast<const T*>(rows + NW * ROW_SLOT);
            const T* sC = reinterpret_cast<const T*>(rows + NW * ROW_SLOT + bc_bytes);
            const float Dv = sc[N], bias = sc[N + 1];
            const LookbackPlan plan = lookback_plan(nt - 1 - c, nt);
            T* gdu = reinterpret_cast<T*>(p.du) + b * p.du_bs + d * p.du_ds + l0;
            T* gdd = reinterpret_cast<T*>(p.ddelta) + b * p.dd_bs + d * p.dd_ds + l0;
            const bool vec_store = !partial && ((reinterpret_cast<uintptr_t>(gdu) | reinterpret_cast<uintptr_t>(gdd)) & 15) == 0;
            float dD_acc = 0.f, dbias_acc = 0.f;

            if constexpr (N1) {
                uint4* aggrow = p.desc + row * nt;
                uint4* inclrow = p.desc_incl + row * nt;
                const uint4* lb_addr = lookback_addr(aggrow, inclrow, 1, c, +1, plan, lane);
                const uint4 lb_first = lookback_prefetch(lb_addr);   // in flight during the local scans
                const float Av = sc[0];
                const float h_in = sc[3];
                float a[ITEMS], h[ITEMS], gl[ITEMS], rp[ITEMS];
                float P = 1.f, Vv = 0.f;
                auto fwd_local = [&](auto tag) {
                    constexpr bool PART = decltype(tag)::value;
#pragma unroll
                    for (int v = 0; v < ITEMS / VT; ++v) {
                        float uv[VT], dl[VT], Bv[VT];
                        lds_items<T, VT>(su + e0 + v * VT, uv);
                        lds_items<T, VT>(sd + e0 + v * VT, dl);
                        lds_items<T, VT>(sB + e0 + v * VT, Bv);
                        if constexpr (kAcc && VT % 2 == 0) {
                            // fp32: softplus, decay and b for two positions at a time on the packed fp32 pipe (scan_common.cuh;
                            // bit-identical to the scalar forms); the recurrence stays scalar
#pragma unroll
                            for (int k = 0; k < VT; k += 2) {
                                f32x2 x2 = add2(pk2(dl[k], dl[k + 1]), splat2(bias));
                                if (softplus_on) {
                                    float x0, x1;
                                    upk2(x2, x0, x1);
                                    x2 = softplus2(x0, x1);
                                }
                                float ee[2], bb[2], xx[2];
                                upk2(decay_m1_2(mul2(x2, splat2(Av))), ee[0], ee[1]);
                                upk2(mul2(mul2(x2, pk2(uv[k], uv[k + 1])), pk2(Bv[k], Bv[k + 1])), bb[0], bb[1]);
                                upk2(x2, xx[0], xx[1]);
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const int i = v * VT + k + j;
                                    if (PART && e0 + i >= len) {
                                        ee[j] = 0.f;
                                        bb[j] = 0.f;
                                        xx[j] = 0.f;
                                    }
                                    dl[k + j] = xx[j];
                                    a[i] = ee[j];
                                    decay_step(ee[j], bb[j], P, Vv);
                                    h[i] = Vv;
                                    rp[i] = P;
                                }
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < VT; ++k) {
                            const int i = v * VT + k;
                            float x = dl[k] + bias;
                            if (softplus_on) x = softplus_f(x);
                            float ei = decay_m1<kAcc>(x * Av);
                            float bi = x * uv[k] * Bv[k];
                            if (PART && e0 + i >= len) {
                                ei = 0.f;
                                bi = 0.f;
                                x = 0.f;
                            }
                            dl[k] = x;
                            a[i] = ei;   // decay minus one
                            decay_step(ei, bi, P, Vv);
                            h[i] = Vv;   // local inclusive state
                            rp[i] = P;   // local inclusive decay (temporarily)
                        }
                        }
                        // keep the activated delta (fp32 accuracy matters only for fp32 inputs) for the output pass:
                        // same lane, same addresses, so no cross-lane hazard
                        if constexpr (sizeof(T) == 4) sts_items<T, VT>(const_cast<T*>(sd) + e0 + v * VT, dl);
                    }
                };
                if (partial) fwd_local(std::true_type{});
                else fwd_local(std::false_type{});
                const float Pth = P;
                warp_scan_fwd(P, Vv, lane);
                float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, Vv, 1);
                if (lane == 0) {
                    Pe = 1.f;
                    Ve = 0.f;
                }
                const float seed = fmaf(Pe, h_in, Ve);
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) h[i] = fmaf(rp[i], seed, h[i]);   // true forward state h_t
                // reverse local scan of g'_t = a_t (C_t dout_t + g'_{t+1})
                float rr = 0.f, RP = 1.f;
                auto rev_local = [&](auto tag) {
                    constexpr bool PART = decltype(tag)::value;
#pragma unroll
                    for (int v = ITEMS / VT - 1; v >= 0; --v) {
                        float Cv[VT], dy[VT];
                        lds_items<T, VT>(sC + e0 + v * VT, Cv);
                        lds_items<DT, VT>(sdo + e0 + v * VT, dy);
#pragma unroll
                        for (int k = VT - 1; k >= 0; --k) {
                            const int i = v * VT + k;
                            const float cd = (PART && e0 + i >= len) ? 0.f : Cv[k] * dy[k];
                            gl[i] = cd + rr;   // g_t with zero incoming adjoint
                            rp[i] = RP;        // d g_t / d incoming
                            rr = fmaf(a[i], gl[i], gl[i]);
                            RP = fmaf(a[i], RP, RP);
                        }
                    }
                };
                if (partial) rev_local(std::true_type{});
                else rev_local(std::false_type{});
                float Pr = Pth, Rr = rr;
                warp_scan_rev(Pr, Rr, lane);
                float Ps = __shfl_down_sync(FULL, Pr, 1), Rs = __shfl_down_sync(FULL, Rr, 1);
                if (lane == 31) {
                    Ps = 1.f;
                    Rs = 0.f;
                }
                const float Pa = __shfl_sync(FULL, Pr, 0), Ra = __shfl_sync(FULL, Rr, 0);
                if (lane == 0 && plan.publish_agg) st_desc(aggrow + c, Pa, Ra, desc_tag(ep, DESC_READY));
                float r_in = 0.f, Psuf = 1.f;
                if (plan.nlanes) {
                    const float2 suf = lookback_finish(lb_addr, lb_first, plan.nlanes, lane, p.err, ep);
                    Psuf = suf.x;
                    r_in = suf.y;
                }
                if (lane == 0 && plan.publish_incl) st_desc(inclrow + c, Pa * Psuf, fmaf(Pa, r_in, Ra), desc_tag(ep, DESC_READY));
                const float rin_t = fmaf(Ps, r_in, Rs);   // adjoint entering this lane's last position
                // outputs, one 128-bit vector of T at a time; dB / dC contributions go to this warp's rows of `red`
                float* accB = red + warp * CL + e0;
                float* accC = red + NW * CL + warp * CL + e0;
                float dA_acc = 0.f;
                auto outputs = [&](auto tag) {
                    constexpr bool PART = decltype(tag)::value;
#pragma unroll
                    for (int v = 0; v < ITEMS / VT; ++v) {
                        float uv[VT], dl[VT], Bv[VT], dy[VT], cB[VT], cC[VT], duv[VT], ddv[VT];
                        lds_items<T, VT>(su + e0 + v * VT, uv);
                        lds_items<T, VT>(sd + e0 + v * VT, dl);   // fp32: already activated (written back above)
                        lds_items<T, VT>(sB + e0 + v * VT, Bv);
                        lds_items<DT, VT>(sdo + e0 + v * VT, dy);
                        if (j > 0) {
                            lds_items<float, VT>(accB + v * VT, cB);
                            lds_items<float, VT>(accC + v * VT, cC);
                        } else {
#pragma unroll
                            for (int k = 0; k < VT; ++k) cB[k] = cC[k] = 0.f;
                        }
                        if constexpr (kAcc && VT % 2 == 0) {
                            // fp32: the ~25 independent fp32 operations of a position, two positions per instruction. Same
                            // operations and association as the scalar branch below.
#pragma unroll
                            for (int k = 0; k < VT; k += 2) {
                                const int i = v * VT + k;
                                const bool v0 = !(PART && e0 + i >= len), v1 = !(PART && e0 + i + 1 >= len);
                                const f32x2 x2 = pk2(dl[k], dl[k + 1]);                       // activated delta (written back above)
                                const f32x2 u2 = pk2(v0 ? uv[k] : 0.f, v1 ? uv[k + 1] : 0.f);
                                const f32x2 dy2 = pk2(v0 ? dy[k] : 0.f, v1 ? dy[k + 1] : 0.f);
                                const f32x2 B2 = pk2(v0 ? Bv[k] : 0.f, v1 ? Bv[k + 1] : 0.f);
                                const f32x2 h2 = pk2(h[i], h[i + 1]);
                                const f32x2 gt2 = fma2(pk2(rp[i], rp[i + 1]), splat2(rin_t), pk2(gl[i], gl[i + 1]));
                                const f32x2 bi2 = mul2(mul2(x2, u2), B2);
                                f32x2 tt2 = mul2(gt2, add2(h2, mul2(bi2, splat2(-1.f))));          // g_t * a_t * h_{t-1} = g_t * (h_t - b_t)
                                if (PART) {
                                    float t0, t1;
                                    upk2(tt2, t0, t1);
                                    tt2 = pk2(v0 ? t0 : 0.f, v1 ? t1 : 0.f);
                                }
                                const f32x2 gx2 = mul2(gt2, x2);
                                const f32x2 du2 = fma2(gx2, B2, mul2(splat2(Dv), dy2));
                                f32x2 ddl2 = fma2(mul2(gt2, u2), B2, mul2(splat2(Av), tt2));
                                float xt0, xt1;
                                upk2(mul2(x2, tt2), xt0, xt1);
                                dA_acc += xt0;                                                  // dA_acc = fma(x, tt, dA_acc): product rounded first here
                                dA_acc += xt1;
                                const f32x2 cB2 = fma2(gx2, u2, pk2(cB[k], cB[k + 1]));
                                const f32x2 cC2 = fma2(dy2, h2, pk2(cC[k], cC[k + 1]));
                                upk2(cB2, cB[k], cB[k + 1]);
                                upk2(cC2, cC[k], cC[k + 1]);
                                float yu0, yu1;
                                upk2(mul2(dy2, u2), yu0, yu1);
                                dD_acc += yu0;
                                dD_acc += yu1;
                                if (softplus_on) ddl2 = mul2(ddl2, mul2(decay_m1_2(mul2(x2, splat2(-1.f))), splat2(-1.f)));   // sigmoid = -expm1(-x)
                                float d0, d1;
                                upk2(ddl2, d0, d1);
                                if (!v0) d0 = 0.f;
                                if (!v1) d1 = 0.f;
                                dbias_acc += d0;
                                dbias_acc += d1;
                                ddv[k] = d0;
                                ddv[k + 1] = d1;
                                upk2(du2, duv[k], duv[k + 1]);
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < VT; ++k) {
                            const int i = v * VT + k;
                            const bool valid = !(PART && e0 + i >= len);
                            float x;
                            if constexpr (sizeof(T) == 4) {
                                x = dl[k];
                            } else {
                                const float raw = valid ? dl[k] + bias : 0.f;   // stale smem beyond the end may hold NaN patterns
                                x = softplus_on ? softplus_f(raw) : raw;
                            }
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
                            // d softplus = sigmoid(raw) = 1 - exp(-softplus(raw)); equals the reference's switch to 1
                            // above raw = 20 (bwd_kernel_oflex.cuh:250-255) to 2e-9
                            if (softplus_on) ddl *= -decay_m1<true>(-x);   // expm1 form: 1 - exp(-x) cancels for the tiny deltas of slow channels
                            if (!valid) ddl = 0.f;
                            dbias_acc += ddl;
                            ddv[k] = ddl;
                        }
                        }
                        sts_items<float, VT>(accB + v * VT, cB);
                        sts_items<float, VT>(accC + v * VT, cC);
                        if (vec_store) {
                            uint4 ru, rd;
                            T* eu = reinterpret_cast<T*>(&ru);
                            T* ed = reinterpret_cast<T*>(&rd);
#pragma unroll
                            for (int k = 0; k < VT; ++k) {
                                eu[k] = ElemTraits<T>::from_f(duv[k]);
                                ed[k] = ElemTraits<T>::from_f(ddv[k]);
                            }
                            reinterpret_cast<uint4*>(gdu + e0)[v] = ru;
                            reinterpret_cast<uint4*>(gdd + e0)[v] = rd;
                        } else {
#pragma unroll
                            for (int k = 0; k < VT; ++k) {
                                const int e = e0 + v * VT + k;
                                if (e < len) {
                                    gdu[e] = ElemTraits<T>::from_f(duv[k]);
                                    gdd[e] = ElemTraits<T>::from_f(ddv[k]);
                                }
                            }
                        }
                    }
                };
                if (partial) outputs(std::true_type{});
                else outputs(std::false_type{});
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dA_acc += __shfl_xor_sync(FULL, dA_acc, o);
                if (lane == 0) atomicAdd(p.dA + d, dA_acc);
            } else {
                // ------------------------------ general dstate ------------------------------
                float du[ITEMS], dd[ITEMS];
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
                // pass 1: tile aggregates of the reverse scan, one state per lane
                float aggP = 1.f, aggR = 0.f;
                for (int n = 0; n < N; ++n) {
                    const float An = sc[n];
                    float Cv[ITEMS];
                    lds_items<T, ITEMS>(sC + n * CL + e0, Cv);
                    float r = 0.f, RP = 1.f;
#pragma unroll
                    for (int i = ITEMS - 1; i >= 0; --i) {
                        const bool valid = !(partial && e0 + i >= len);
                        const float ei = valid ? decay_m1<kAcc>(dl[i] * An) : 0.f;
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
                uint4* aggrow = p.desc + (row * nt) * N;
                uint4* inclrow = p.desc_incl + (row * nt) * N;
                if (lane < N && plan.publish_agg) st_desc(aggrow + (int64_t)c * N + lane, aggP, aggR, desc_tag(ep, DESC_READY));
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
                                const float2 suf = lookback_finish(addr[q], first[q], plan.nlanes, lane, p.err, ep);
                                if (lane == n0 + q) {
                                    sufP = suf.x;
                                    sufR = suf.y;
                                }
                            }
                        }
                    }
                }
                if (lane < N && plan.publish_incl) st_desc(inclrow + (int64_t)c * N + lane, aggP * sufP, fmaf(aggP, sufR, aggR), desc_tag(ep, DESC_READY));
                // pass 2: per state, forward states from the carry, reverse adjoints from the look-back
                for (int n = 0; n < N; ++n) {
                    const float Av = sc[n];
                    float a[ITEMS], h[ITEMS], gl[ITEMS], rp[ITEMS], Bv[ITEMS];
                    lds_items<T, ITEMS>(sB + n * CL + e0, Bv);
                    float P = 1.f, V = 0.f;
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const bool valid = !(partial && e0 + i >= len);
                        const float ei = valid ? decay_m1<kAcc>(dl[i] * Av) : 0.f;
                        const float bi = valid ? dl[i] * uv[i] * Bv[i] : 0.f;
                        a[i] = ei;   // decay minus one
                        decay_step(ei, bi, P, V);
                        h[i] = V;
                        rp[i] = P;
                    }
                    const float Pth = P;
                    warp_scan_fwd(P, V, lane);
                    float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
                    if (lane == 0) {
                        Pe = 1.f;
                        Ve = 0.f;
                    }
                    const float seed = fmaf(Pe, sc[N + 2 + n], Ve);
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
                float raw[ITEMS];
                lds_items<T, ITEMS>(sd + e0, raw);
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const bool valid = !(partial && e0 + i >= len);
                    float ddl = dd[i];
                    if (p.softplus) ddl *= softplus_grad_f(valid ? raw[i] + bias : 0.f);
                    if (!valid) ddl = 0.f;
                    dd[i] = ddl;
                    dbias_acc += ddl;
                }
                if (vec_store) {
#pragma unroll
                    for (int v = 0; v < ITEMS / VT; ++v) {
                        uint4 ru, rd;
                        T* eu = reinterpret_cast<T*>(&ru);
                        T* ed = reinterpret_cast<T*>(&rd);
#pragma unroll
                        for (int k = 0; k < VT; ++k) {
                            eu[k] = ElemTraits<T>::from_f(du[v * VT + k]);
                            ed[k] = ElemTraits<T>::from_f(dd[v * VT + k]);
                        }
                        reinterpret_cast<uint4*>(gdu + e0)[v] = ru;
                        reinterpret_cast<uint4*>(gdd + e0)[v] = rd;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const int e = e0 + i;
                        if (e < len) {
                            gdu[e] = ElemTraits<T>::from_f(du[i]);
                            gdd[e] = ElemTraits<T>::from_f(dd[i]);
                        }
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
        } else {
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
        __syncwarp();   // every lane is done reading the stage
        if (lane == 0) mbar_arrive(&empty[s]);

        if constexpr (N1) {
            if (tc.aux1) {
                // ---------------- flush dB / dC of this (b, g, split, chunk) slab ----------------
                float* redB = red;
                float* redC = red + NW * CL;
                named_bar_sync(1, NW * 32);
                constexpr int PER_WARP = CL / NW;
                float* gB = p.dB + ((int64_t)b * p.G + g) * p.L + l0;
                float* gC = p.dC + ((int64_t)b * p.G + g) * p.L + l0;
                for (int e = warp * PER_WARP + lane; e < (warp + 1) * PER_WARP; e += 32) {
                    float sb = 0.f, scc = 0.f;
#pragma unroll
                    for (int r = 0; r < NW; ++r) {
                        sb += redB[r * CL + e];
                        scc += redC[r * CL + e];
                    }
                    if (e < len) {
                        if (p.atomic_bc) {
                            atomicAdd(gB + e, sb);
                            atomicAdd(gC + e, scc);
                        } else {
                            gB[e] = sb;
                            gC[e] = scc;
                        }
                    }
                }
                named_bar_sync(1, NW * 32);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, typename DT, int ITEMS, bool N1>
static int launch_bwd(ScanBwdArgs& a, int sm_count, cudaStream_t stream) {
    constexpr int NW = kScanWarps;
    constexpr int CL = 32 * ITEMS;
    auto kernel = scan_bwd_kernel<T, DT, ITEMS, NW, N1, -1>;
    if constexpr (N1 && sizeof(T) == 4 && sizeof(DT) == 4)      // the BEM configuration: softplus resolved at compile time
        kernel = a.softplus ? scan_bwd_kernel<T, DT, ITEMS, NW, N1, 1> : scan_bwd_kernel<T, DT, ITEMS, NW, N1, 0>;
    const int hdr_bytes = 128 + ((NW * (2 * a.N + 2) * 4 + 127) / 128) * 128;
    const int stage_bytes = hdr_bytes + NW * (2 * CL * (int)sizeof(T) + CL * (int)sizeof(DT)) + 2 * a.N * CL * (int)sizeof(T);
    const int red_bytes = N1 ? 2 * NW * CL * (int)sizeof(float) : 0;
    const int fixed = red_bytes + 512;
    const int budget2 = (227 * 1024) / 2 - 1024;
    const bool two_ok = N1 && sizeof(T) == 4;   // matches __launch_bounds__ of the kernel
    int stages, ctas_per_sm;
    if (two_ok && 2 * stage_bytes + fixed <= budget2) {
        stages = min(3, (budget2 - fixed) / stage_bytes);
        ctas_per_sm = 2;
    } else {
        stages = min(4, (227 * 1024 - fixed) / stage_bytes);
        ctas_per_sm = 1;
        if (stages < 2) return BEM_ERR_UNSUPPORTED;
    }
    a.stages = stages;
    const int smem_bytes = stages * stage_bytes + red_bytes + stages * 2 * 8 + 64;
    static int cached_smem_v[2][64] = {{0}};      // per kernel function: [softplus variant][device]
    int* cached_smem = cached_smem_v[a.softplus ? 1 : 0];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] != smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        cached_smem[dev] = smem_bytes;
    }
    const int64_t ndesc = (int64_t)a.batch * a.dim * a.nchunks * a.N;
    a.desc_incl = a.desc + ndesc;
    const int grid = min(a.total_tiles, sm_count * ctas_per_sm);
    launch_pdl(kernel, dim3(grid), dim3((NW + 1) * 32), smem_bytes, stream, a);
    return (int)cudaGetLastError();
}

template <typename T, typename DT, int ITEMS>
static int launch_bwd_n(ScanBwdArgs& a, int sm_count, cudaStream_t stream) {
    if (a.N == 1) return launch_bwd<T, DT, ITEMS, true>(a, sm_count, stream);
    return launch_bwd<T, DT, ITEMS, false>(a, sm_count, stream);
}

int scan_bwd_dispatch(ScanBwdArgs& a, int dtype, int dout_dtype, int sm_count, cudaStream_t stream) {
    if (scan_rows_preferred(a.batch, a.dim, a.N, sm_count)) return scan_rows_bwd_dispatch(a, dtype, dout_dtype, sm_count, stream);
    if (a.N > kMaxDstate) return BEM_ERR_UNSUPPORTED;
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
