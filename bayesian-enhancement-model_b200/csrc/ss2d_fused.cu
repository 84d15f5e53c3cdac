// ss2d_fused.cu — the SS2D core as a traversal-aware scan: cross-scan gather, four-direction selective scan (dt_proj fused)
// and cross-merge scatter in ONE pass over the image, no materialised traversal.
//
// What it replaces: the part of SS2Dv2.forward_corev2 after x_proj (basicsr/vmamba/models/vmamba.py:657-684):
//     xs = cross_scan(x)                     (B, 4, D, L)   csm_triton.py:22-34, 278-390
//     dts = dt_proj(dt), split B / C         (B, 4*D, L)    vmamba.py:660-661
//     ys = selective_scan(xs, dts, A, B, C, D, bias, softplus)      csms6s.py:116-130, cusoflex/*_fwd_kernel_oflex.cuh
//     y  = cross_merge(ys) = (y0 + flip(y2)) + T(y1 + flip(y3))     csm_triton.py:60-62
// The first version of bem_ss2d_fwd composed four launches and moved ~770 MB per level-0 core (600x400, D = 40).
//
// Idea. The four traversals are the row-major walk of the image (k0), the column-major walk (k1) and their reversals (k2, k3).
// Scanning a reversed sequence forward is scanning the sequence backward in place, so every direction can read the image where
// it lies and leave its result at the pixel it belongs to. What is left of the traversals is WHICH neighbour the recurrence
// comes from: left / right for k0 / k2, up / down for k1 / k3 (with the wrap from the end of one image row / column to the start
// of the next). A 32 x 32 pixel tile therefore contains 32 row segments and 32 column segments of each channel's sequences,
// and the scan splits into the classic three steps, all in image coordinates:
//   1. ss2d_tile_kernel<R, false>: per tile, per channel, per direction: the affine map (P, V) of every segment
//      (h_out = P h_in + V), 1.2 M maps for the level-0 shape;
//   2. ss2d_carry_kernel: per (channel, direction) the exclusive scan of its segments' maps in flow order -> the state entering
//      every segment (7600 / 7800 segments per sequence: row-major order of (h, tile column) resp. column-major (w, tile row));
//   3. ss2d_tile_kernel<R, true>: per tile again, every segment re-walked from its true incoming state; the four directions of a
//      pixel are summed on chip in the reference's association (y0 + y2) + (y1 + y3) and y is written once, coalesced.
// Both tile passes read x and the image-order x_proj output with coalesced row loads into a padded shared tile (pitch 33): the
// row walks (lane = image row) and the column walks (lane = image column) are both free of bank conflicts. No transposed copy,
// no flip, no (B, 4, D, L) tensor. DRAM traffic per core: x twice, xdbl twice, y once + 12 B per segment ~= 190 MB at level 0.
// The price is that delta / decay are evaluated twice (the kernel is bound by instruction issue, not by HBM — as the plain scan
// is); it still halves the time of the composed form, which evaluated them once but moved 4x the data through three more launches.
//
// Numerics: same element arithmetic as scan_fwd*.cu (softplus_f, decay_m1<true>, h <- fma(e, h, h) + b). Segment maps are
// composed in fp32 in a fixed order (bit-reproducible, independent of the grid).
#include <cstdlib>
#include <type_traits>

#include "bem_kernels.h"
#include "scan_common.cuh"

namespace bem {

namespace {
constexpr int TS = 32;            // tile edge = segment length
constexpr int PITCH = TS + 1;     // padded shared pitch: conflict-free for lane = row and for lane = column
constexpr int DB = 8;             // channels per CTA (one per warp)
constexpr int THREADS = DB * 32;

struct TileGeom {
    int b, d0, h0, w0, th, tw, ti, tj;
};

// segment arrays (agg, hin): [k0 : BD * NSr][k2 : BD * NSr][k1 : BD * NSc][k3 : BD * NSc], memory order inside a sequence
__device__ __forceinline__ int64_t seg_offset(const Ss2dFusedArgs& p, int k) {
    const int64_t BD = (int64_t)p.B * p.D;
    const int64_t NSr = (int64_t)p.H * p.NTW, NSc = (int64_t)p.W * p.NTH;
    return ((k & 1) ? 2 * BD * NSr : 0) + ((k >> 1) ? BD * ((k & 1) ? NSc : NSr) : 0);
}

__device__ __forceinline__ TileGeom tile_geom(const Ss2dFusedArgs& p) {
    TileGeom g;
    const int nDb = (p.D + DB - 1) / DB;
    g.tj = blockIdx.x;
    g.ti = blockIdx.y;
    g.b = blockIdx.z / nDb;
    g.d0 = (blockIdx.z - g.b * nDb) * DB;
    g.h0 = g.ti * TS;
    g.w0 = g.tj * TS;
    g.th = min(TS, p.H - g.h0);
    g.tw = min(TS, p.W - g.w0);
    return g;
}
}  // namespace

// One 32 x 32 tile x 8 channels. APPLY = false: segment maps; APPLY = true: outputs.
// Row walks: warp = channel, lane = image row of the tile; column walks: warp = channel, lane = image column.
template <int R, bool APPLY, bool SOFTPLUS>
__global__ void __launch_bounds__(THREADS, APPLY ? 2 : 3) ss2d_tile_kernel(const Ss2dFusedArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr int CX = R + 2;                       // [dt rows | B | C] of one direction (dstate 1)
    // projected channels staged per direction PAIR (two loads per tile) while that keeps two CTAs on an SM (dt_rank 3: 110 KB),
    // else per direction (four loads, half the buffer): dt_rank 5 would otherwise run one CTA = 8 warps per SM
    constexpr bool PAIR = R <= 3;
    constexpr int XD_DIRS = PAIR ? 2 : 1;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                                 // [DB][TS][PITCH]   u
    float* acc = xs + DB * TS * PITCH;              // [DB][TS][PITCH]   y0 + y2 (APPLY only)
    float* xd = APPLY ? acc + DB * TS * PITCH : acc;   // [XD_DIRS][CX][TS][PITCH] projected channels of the staged direction(s)
    const TileGeom g = tile_geom(p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = g.d0 + warp;
    const bool dval = d < p.D;
    const int64_t HW = (int64_t)p.H * p.W;

    // Tile loads are 4-byte cp.async copies (zero-filled outside the image): a warp issues one image row of 32 pixels per
    // instruction (coalesced 128 B) and all of a thread's copies are in flight together (a load -> store loop serialises ~40
    // DRAM round trips per warp).
    auto cp4 = [](float* dst, const float* src, bool valid) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 4u : 0u) : "memory");
    };
    // Addresses advance by pointer increments and an out-of-image copy reads zero bytes from the (clamped, always valid) address
    // of the tile's first pixel column / row, so a copy costs ~4 instructions; the loads are ~15 % of the kernel's instructions.
    const int lane_c = min(lane, g.tw - 1);
    const uint32_t col_ok = lane < g.tw ? 4u : 0u;
    auto cp_rows = [&](float* dst, const float* src, int row0, int row_step, int n_rows) {
        // rows row0, row0 + row_step, ... of a (TS x TS) tile plane whose pixel (0, lane_c) is at `src`; dst likewise
        uint32_t sdst = smem_u32(dst) + (uint32_t)(row0 * PITCH + lane) * 4u;
        const float* gsrc = src + (int64_t)min(row0, g.th - 1) * p.W;
        const int64_t gstep = (int64_t)row_step * p.W;
        int h = row0;
#pragma unroll 4
        for (int i = 0; i < n_rows; ++i) {
            const bool ok = h < g.th;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sdst), "l"(gsrc), "r"(ok ? col_ok : 0u) : "memory");
            sdst += (uint32_t)(row_step * PITCH) * 4u;
            h += row_step;
            if (h < g.th) gsrc += gstep;          // stays inside the image
        }
    };
    // ---- u tile: warp `warp` loads its channel ----
    cp_rows(xs + warp * TS * PITCH, p.x + ((int64_t)g.b * p.D + (dval ? d : 0)) * HW + (int64_t)g.h0 * p.W + g.w0 + lane_c, 0, 1,
            dval ? TS : 0);
    if (!dval) {   // a channel slot past D: zeros, so that its lanes compute on defined data
        for (int i = lane; i < TS * PITCH; i += 32) xs[warp * TS * PITCH + i] = 0.f;
    }
    // projected channels of directions pp, pp + 2: 2 * CX planes of TS rows, four rows per warp and plane
    auto load_xd = [&](int k0, int ndirs) {                    // directions k0, k0 + 2, ... (ndirs of them) into xd
        const float* plane = p.xdbl + ((int64_t)g.b * 4 + k0) * CX * HW + (int64_t)g.h0 * p.W + g.w0 + lane_c;
#pragma unroll 1
        for (int qc = 0; qc < ndirs * CX; ++qc) {
            cp_rows(xd + qc * TS * PITCH, plane, warp, DB, TS / DB);
            plane += (qc == CX - 1) ? (int64_t)(CX + 1) * HW : HW;      // direction k0 + 2 starts 2 * CX planes after direction k0
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    };
    load_xd(0, XD_DIRS);
    __syncthreads();

    const int64_t NSr = (int64_t)p.H * p.NTW, NSc = (int64_t)p.W * p.NTH;     // segments per (b, d) sequence: rows / columns
    const int64_t bd = (int64_t)g.b * p.D + d;
    constexpr int PL = TS * PITCH;                                            // one shared plane

    // ------------------------------------------------------------------------------------------------------------------
    // One direction of one pair. ROWS: this lane owns image row `lane` of the tile and walks its pixels; else it owns image
    // column `lane` and walks down / up. q = 0 forward (k0 / k1), q = 1 backward (k2 / k3). The walks are ROLLED loops (4 steps
    // per iteration): fully unrolled they are 130 KB of straight-line code that every CTA executes exactly once — the first
    // version spent its time in instruction-cache misses (a constant ~200 us whatever the image size).
    // Where the partial sums live: k0 writes y0 to `acc`, k2 adds y2 (acc = y0 + y2); k1 parks y1 in the output buffer itself
    // (same thread writes and re-reads it, L2-resident), k3 forms (y0 + y2) + (y1 + y3) and stores the final value.
    // ------------------------------------------------------------------------------------------------------------------
    auto walk = [&](auto rows_tag, auto q_tag) {
        constexpr bool ROWS = decltype(rows_tag)::value;
        constexpr int q = decltype(q_tag)::value;
        const int k = (ROWS ? 0 : 1) + 2 * q;
        const int kd = k * p.D + (dval ? d : 0);
        const float A1 = p.A[kd];
        const float Dv = p.Ds ? p.Ds[kd] : 0.f;
        const float bias = p.bias ? p.bias[kd] : 0.f;
        float wdt[R];
#pragma unroll
        for (int r = 0; r < R; ++r) wdt[r] = p.dt_w[(int64_t)kd * R + r];
        const int n_own = ROWS ? g.th : g.tw;           // lanes that own a real row / column
        const int n_step = ROWS ? g.tw : g.th;          // real pixels along the walk (the rest of the tile is outside the image)
        const bool own = dval && lane < n_own;
        // segment index of this lane's row / column segment in its sequence (memory order)
        const int64_t seg = ROWS ? ((int64_t)(g.h0 + lane) * p.NTW + g.tj) : ((int64_t)(g.w0 + lane) * p.NTH + g.ti);
        const int64_t NS = ROWS ? NSr : NSc;
        const int64_t slot = seg_offset(p, k) + bd * NS + seg;
        float h = 0.f, P = 1.f;
        if (APPLY && own) h = p.hin[slot];
        constexpr int ustep = ROWS ? 1 : PITCH;
        // APPLY walks in flow order. The map pass walks AGAINST the flow: with T = sum of the deltas between a pixel and the
        // segment's flow end, V = sum_t exp(A T_t) b_t and P = exp(A T_total) — no recurrence, one ex2 per pixel and no expm1
        // polynomial (an ex2.approx error enters V additively, 2.4e-7 relative, instead of compounding through a product of decays;
        // P comes from ONE accurate expm1 of the exact sum). 31 instead of ~50 instructions per pixel.
        constexpr bool ASC = APPLY ? (q == 0) : (q == 1);
        constexpr int step = ASC ? ustep : -ustep;
        const int first = ASC ? 0 : (n_step - 1);                       // first visited pixel along the walk
        const int lane_off = ROWS ? lane * PITCH : lane;
        const float* us = xs + warp * PL + lane_off + first * ustep;
        const float* xq = xd + (PAIR ? q : 0) * CX * PL + lane_off + first * ustep;
        float* ac = acc + warp * PL + lane_off + first * ustep;
        // column walks: this lane's pixel column in the output image
        float* gy = p.y + ((int64_t)g.b * p.D + (dval ? d : 0)) * HW + (int64_t)(g.h0 + first) * p.W + g.w0 + lane;
        const int64_t gstep = ASC ? p.W : -(int64_t)p.W;
        const float A2 = A1 * kLog2e;
        float T = 0.f;
        (void)ac;
        (void)gy;
        (void)gstep;
        (void)Dv;
        (void)P;
        (void)A2;
        (void)T;
        (void)gstep;
        // one pixel; y1v: the parked y1 of this pixel (k3 only)
        auto pixel = [&](float y1v) {
            const float u = us[0];
            float dl = bias;
#pragma unroll
            for (int r = 0; r < R; ++r) dl = fmaf(wdt[r], xq[r * PL], dl);
            const float Bv = xq[R * PL];
            if constexpr (SOFTPLUS) dl = softplus_f(dl);
            const float bb = dl * u * Bv;
            if constexpr (APPLY) {
                const float e = decay_m1<true>(dl * A1);
                h = fmaf(e, h, h) + bb;
                const float yv = fmaf(xq[(R + 1) * PL], h, Dv * u);
                if constexpr (ROWS) {
                    ac[0] = q == 0 ? yv : ac[0] + yv;                       // y0, then y0 + y2
                } else if constexpr (q == 0) {
                    if (own) gy[0] = yv;                                    // y1 parked where the result will go
                } else {
                    if (own) gy[0] = ac[0] + (y1v + yv);                    // (y0 + y2) + (y1 + y3)
                }
                gy += gstep;
                ac += step;
            } else {
                h = fmaf(ex2_approx(A2 * T), bb, h);                        // h doubles as V
                T += dl;
            }
            us += step;
            xq += step;
        };
        // two consecutive pixels of the walk: everything that does not depend on the recurrence (dt_proj, softplus, decay, b,
        // D u) on the packed fp32 pipe (scan_common.cuh: FFMA2 / FMUL2 / FADD2, bit-identical to the scalar forms), then
        // the two recurrence steps
        auto pixel_pair = [&](float y1a, float y1b) {
            const f32x2 u2 = pk2(us[0], us[step]);
            f32x2 dl2 = splat2(bias);
#pragma unroll
            for (int r = 0; r < R; ++r) dl2 = fma2(splat2(wdt[r]), pk2(xq[r * PL], xq[r * PL + step]), dl2);
            const f32x2 B2 = pk2(xq[R * PL], xq[R * PL + step]);
            float d0, d1;
            upk2(dl2, d0, d1);
            if constexpr (SOFTPLUS) {
                dl2 = softplus2(d0, d1);
                upk2(dl2, d0, d1);
            }
            float bb[2];
            upk2(mul2(mul2(dl2, u2), B2), bb[0], bb[1]);
            if constexpr (APPLY) {
                float e[2], du[2];
                upk2(decay_m1_2(mul2(dl2, splat2(A1))), e[0], e[1]);
                upk2(mul2(splat2(Dv), u2), du[0], du[1]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    h = fmaf(e[j], h, h) + bb[j];
                    const float yv = fmaf(xq[(R + 1) * PL + j * step], h, du[j]);
                    if constexpr (ROWS) {
                        ac[0] = q == 0 ? yv : ac[0] + yv;
                    } else if constexpr (q == 0) {
                        if (own) gy[0] = yv;
                    } else {
                        if (own) gy[0] = ac[0] + ((j == 0 ? y1a : y1b) + yv);
                    }
                    gy += gstep;
                    ac += step;
                }
            } else {
                h = fmaf(ex2_approx(A2 * T), bb[0], h);
                T += d0;
                h = fmaf(ex2_approx(A2 * T), bb[1], h);
                T += d1;
            }
            us += 2 * step;
            xq += 2 * step;
        };
        constexpr bool K3 = APPLY && !ROWS && q == 1;
        constexpr int CH = APPLY ? 8 : 4;                                   // pixels per loop iteration
        // k3 re-reads the y1 values k1 parked in global memory: the loads of the NEXT chunk are issued before this chunk's
        // arithmetic (the compiler cannot hoist them itself across the chunk's stores to the same array)
        float ynext[CH];
        auto fetch = [&](int s0) {
#pragma unroll
            for (int j = 0; j < CH; ++j) ynext[j] = (K3 && own && s0 + j < n_step) ? gy[(int64_t)j * gstep] : 0.f;
        };
        if constexpr (K3) fetch(0);
        int s = 0;
        for (; s + CH <= n_step; s += CH) {
            float ycur[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) ycur[j] = ynext[j];
            if constexpr (K3) {
                const float* keep = gy;
                gy += (int64_t)CH * gstep;
                fetch(s + CH);
                gy = const_cast<float*>(keep);
            }
#pragma unroll
            for (int j = 0; j < CH; j += 2) pixel_pair(ycur[j], ycur[j + 1]);
        }
        for (int j = 0; s < n_step; ++s, ++j) {                             // ragged end of an edge tile
            float y1v = 0.f;
            if constexpr (K3) y1v = own ? gy[0] : 0.f;
            pixel(y1v);
        }
        if constexpr (!APPLY) {
            if (own) p.agg[slot] = make_float2(1.f + decay_m1<true>(A1 * T), h);
        }
    };

    // ---- directions 0 / 2: along image rows ----
    walk(std::true_type{}, std::integral_constant<int, 0>{});
    if constexpr (!PAIR) {
        __syncthreads();
        load_xd(2, 1);
        __syncthreads();
    }
    walk(std::true_type{}, std::integral_constant<int, 1>{});
    __syncthreads();            // everyone is done with the staged channels (acc is written and read by the same warp)
    load_xd(1, XD_DIRS);
    __syncthreads();
    // ---- directions 1 / 3: along image columns ----
    walk(std::false_type{}, std::integral_constant<int, 0>{});
    if constexpr (!PAIR) {
        __syncthreads();
        load_xd(3, 1);
        __syncthreads();
    }
    walk(std::false_type{}, std::integral_constant<int, 1>{});
}

// Exclusive scan of one sequence's segment maps in flow order -> the state entering each segment.
// grid (B * D, 4): blockIdx.y = direction. A round handles 256 x CPT consecutive flow positions: coalesced loads of CPT maps
// per thread all in flight, transposed through shared memory so that each thread composes CPT CONSECUTIVE maps in registers,
// one block-level scan of the 256 thread totals, then the CPT incoming states go back the same way. (The first version walked
// 256 segments per round with two barriers and a dependent load each: 30 serial rounds = 33 us for the level-0 image.)
constexpr int CPT = 16;
__global__ void __launch_bounds__(256) ss2d_carry_kernel(const Ss2dFusedArgs p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float2 buf[256 * CPT + 256 * CPT / 16];      // padded: index i lives at i + i / 16
    __shared__ float2 wtot[8];
    const int k = blockIdx.y;
    const int64_t bd = blockIdx.x;
    const int64_t NSr = (int64_t)p.H * p.NTW, NSc = (int64_t)p.W * p.NTH;
    const int64_t NS = (k & 1) ? NSc : NSr;
    const int64_t dir_off = seg_offset(p, k);
    const float2* agg = p.agg + dir_off + bd * NS;
    float* hin = p.hin + dir_off + bd * NS;
    const bool rev = k >= 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float carry = 0.f;                                   // state entering the current round
    for (int64_t f0 = 0; f0 < NS; f0 += 256 * CPT) {
        float2 m[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {                  // coalesced: thread t takes flow positions f0 + j * 256 + t
            const int64_t f = f0 + j * 256 + tid;
            m[j] = f < NS ? agg[rev ? NS - 1 - f : f] : make_float2(1.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int i = j * 256 + tid;
            buf[i + (i >> 4)] = m[j];
        }
        __syncthreads();
        float P = 1.f, V = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {                  // this thread's CPT consecutive maps, flow order
            const int i = tid * CPT + j;
            m[j] = buf[i + (i >> 4)];
            V = fmaf(m[j].x, V, m[j].y);
            P *= m[j].x;
        }
        warp_scan_fwd(P, V, lane);
        if (lane == 31) wtot[warp] = make_float2(P, V);
        __syncthreads();
        float hw = carry, c = carry;
        for (int w2 = 0; w2 < 8; ++w2) {
            const float2 t = wtot[w2];
            if (w2 < warp) hw = fmaf(t.x, hw, t.y);
            c = fmaf(t.x, c, t.y);
        }
        float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
        if (lane == 0) {
            Pe = 1.f;
            Ve = 0.f;
        }
        float hcur = fmaf(Pe, hw, Ve);                   // state entering this thread's first segment
        float* hbuf = reinterpret_cast<float*>(buf);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int i = tid * CPT + j;
            const float2 mj = m[j];
            __syncwarp();
            hbuf[2 * (i + (i >> 4))] = hcur;             // reuse the slot of map i (x component)
            hcur = fmaf(mj.x, hcur, mj.y);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int i = j * 256 + tid;
            const int64_t f = f0 + i;
            if (f < NS) hin[rev ? NS - 1 - f : f] = hbuf[2 * (i + (i >> 4))];
        }
        carry = c;
        __syncthreads();
    }
}

template <int R, bool SP>
static int launch_fused(const Ss2dFusedArgs& a, cudaStream_t stream) {
    constexpr int CX = R + 2;
    constexpr int XD_DIRS = R <= 3 ? 2 : 1;
    const size_t sm1 = (size_t)(DB * TS * PITCH + XD_DIRS * CX * TS * PITCH) * 4;
    const size_t sm3 = sm1 + (size_t)DB * TS * PITCH * 4;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ss2d_tile_kernel<R, false, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(ss2d_tile_kernel<R, true, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
        if (e != cudaSuccess) return (int)e;
        attr_done[dev] = true;
    }
    const int nDb = (a.D + DB - 1) / DB;
    const dim3 grid(a.NTW, a.NTH, a.B * nDb);
    if ((int64_t)a.B * nDb > 65535 || a.NTH > 65535) return BEM_ERR_UNSUPPORTED;
    launch_pdl(ss2d_tile_kernel<R, false, SP>, grid, dim3(THREADS), sm1, stream, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    launch_pdl(ss2d_carry_kernel, dim3(a.B * a.D, 4), dim3(256), 0, stream, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    launch_pdl(ss2d_tile_kernel<R, true, SP>, grid, dim3(THREADS), sm3, stream, a);
    return (int)cudaGetLastError();
}

bool ss2d_fused_supported(int dstate, int dt_rank) { return dstate == 1 && (dt_rank == 3 || dt_rank == 5 || dt_rank == 10); }

int64_t ss2d_fused_workspace(int B, int D, int H, int W) {
    const int64_t NTH = (H + TS - 1) / TS, NTW = (W + TS - 1) / TS;
    const int64_t nseg = 2 * (int64_t)B * D * ((int64_t)H * NTW + (int64_t)W * NTH);
    return (nseg * 8 + 255) / 256 * 256 + (nseg * 4 + 255) / 256 * 256;
}

int ss2d_fused_dispatch(Ss2dFusedArgs a, void* workspace, cudaStream_t stream) {
    a.NTH = (a.H + TS - 1) / TS;
    a.NTW = (a.W + TS - 1) / TS;
    const int64_t nseg = 2 * (int64_t)a.B * a.D * ((int64_t)a.H * a.NTW + (int64_t)a.W * a.NTH);
    a.agg = reinterpret_cast<float2*>(workspace);
    a.hin = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (nseg * 8 + 255) / 256 * 256);
    switch (a.R) {
        case 3: return a.softplus ? launch_fused<3, true>(a, stream) : launch_fused<3, false>(a, stream);
        case 5: return a.softplus ? launch_fused<5, true>(a, stream) : launch_fused<5, false>(a, stream);
        case 10: return a.softplus ? launch_fused<10, true>(a, stream) : launch_fused<10, false>(a, stream);
        default: return BEM_ERR_UNSUPPORTED;
    }
}

}  // namespace bem
