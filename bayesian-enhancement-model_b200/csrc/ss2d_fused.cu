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
template <int R, bool APPLY>
__global__ void __launch_bounds__(THREADS, 2) ss2d_tile_kernel(const Ss2dFusedArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr int CX = R + 2;                       // [dt rows | B | C] of one direction (dstate 1)
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                                 // [DB][TS][PITCH]   u
    float* acc = xs + DB * TS * PITCH;              // [DB][TS][PITCH]   y0 + y2 (APPLY only)
    float* xd = APPLY ? acc + DB * TS * PITCH : acc;   // [2][CX][TS][PITCH] projected channels of the two directions of a pair
    const TileGeom g = tile_geom(p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = g.d0 + warp;
    const bool dval = d < p.D;
    const int64_t HW = (int64_t)p.H * p.W;

    // ---- u tile: warp `warp` loads its channel, one image row per instruction (coalesced 128 B) ----
    {
        const float* src = p.x + ((int64_t)g.b * p.D + (dval ? d : 0)) * HW + (int64_t)g.h0 * p.W + g.w0 + lane;
        float* dst = xs + warp * TS * PITCH + lane;
        const bool cv = dval && lane < g.tw;
#pragma unroll 8
        for (int h = 0; h < TS; ++h) dst[h * PITCH] = (cv && h < g.th) ? src[(int64_t)h * p.W] : 0.f;
    }
    // projected channels of directions pp, pp + 2: 2 * CX * TS rows of 32 pixels, spread over the 8 warps
    auto load_xd = [&](int pp) {
        for (int row = warp; row < 2 * CX * TS; row += DB) {
            const int q = row / (CX * TS), rem = row - q * (CX * TS);
            const int c = rem / TS, h = rem - c * TS;
            const int k = pp + 2 * q;
            const float* src = p.xdbl + (((int64_t)g.b * 4 + k) * CX + c) * HW + (int64_t)(g.h0 + h) * p.W + g.w0 + lane;
            xd[(q * CX + c) * TS * PITCH + h * PITCH + lane] = (h < g.th && lane < g.tw) ? *src : 0.f;
        }
    };
    load_xd(0);
    __syncthreads();

    const int64_t NSr = (int64_t)p.H * p.NTW, NSc = (int64_t)p.W * p.NTH;     // segments per (b, d) sequence: rows / columns
    const int64_t bd = (int64_t)g.b * p.D + d;

    float yreg[TS];
    // ------------------------------------------------------------------------------------------------------------------
    // One direction of one pair. ROWS: this lane owns image row `lane` of the tile and walks its 32 pixels; else it owns image
    // column `lane` and walks down / up. q = 0 forward (k0 / k1), q = 1 backward (k2 / k3).
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
        const int n_step = ROWS ? g.tw : g.th;          // real pixels along the walk
        const bool own = dval && lane < n_own;
        // segment index of this lane's row / column segment in its sequence (memory order)
        const int64_t seg = ROWS ? ((int64_t)(g.h0 + lane) * p.NTW + g.tj) : ((int64_t)(g.w0 + lane) * p.NTH + g.ti);
        const int64_t NS = ROWS ? NSr : NSc;
        const int64_t slot = seg_offset(p, k) + bd * NS + seg;
        float h = 0.f, P = 1.f;
        if (APPLY && own) h = p.hin[slot];
        const float* us = ROWS ? xs + (warp * TS + lane) * PITCH : xs + warp * TS * PITCH + lane;
        const float* xq = ROWS ? xd + q * CX * TS * PITCH + lane * PITCH : xd + q * CX * TS * PITCH + lane;
        constexpr int ustep = ROWS ? 1 : PITCH;
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            const int i = q == 0 ? s : TS - 1 - s;          // position along the walk, in image coordinates
            const float u = us[i * ustep];
            float dl = bias;
#pragma unroll
            for (int r = 0; r < R; ++r) dl = fmaf(wdt[r], xq[r * TS * PITCH + i * ustep], dl);
            const float Bv = xq[R * TS * PITCH + i * ustep];
            if (p.softplus) dl = softplus_f(dl);
            if (i >= n_step) dl = 0.f;                      // outside the image: identity map (e = 0, b = 0)
            const float e = decay_m1<true>(dl * A1);
            const float bb = dl * u * Bv;
            h = fmaf(e, h, h) + bb;
            if constexpr (APPLY) {
                const float Cv = xq[(R + 1) * TS * PITCH + i * ustep];
                const float yv = fmaf(Cv, h, Dv * u);
                yreg[i] = q == 0 ? yv : yreg[i] + yv;       // (y0 + y2) resp. (y1 + y3)
            } else {
                P = fmaf(e, P, P);
            }
        }
        if constexpr (!APPLY) {
            if (own) p.agg[slot] = make_float2(P, h);
        }
    };

    // ---- directions 0 / 2: along image rows ----
    walk(std::true_type{}, std::integral_constant<int, 0>{});
    walk(std::true_type{}, std::integral_constant<int, 1>{});
    if constexpr (APPLY) {
        float* a = acc + (warp * TS + lane) * PITCH;
#pragma unroll
        for (int i = 0; i < TS; ++i) a[i] = yreg[i];
    }
    __syncthreads();            // everyone is done with the pair-0 channels
    load_xd(1);
    __syncthreads();
    // ---- directions 1 / 3: along image columns ----
    walk(std::false_type{}, std::integral_constant<int, 0>{});
    walk(std::false_type{}, std::integral_constant<int, 1>{});
    if constexpr (APPLY) {
        if (dval && lane < g.tw) {
            float* dst = p.y + ((int64_t)g.b * p.D + d) * HW + (int64_t)g.h0 * p.W + g.w0 + lane;
            const float* a = acc + warp * TS * PITCH + lane;
#pragma unroll
            for (int i = 0; i < TS; ++i)
                if (i < g.th) dst[(int64_t)i * p.W] = a[i * PITCH] + yreg[i];      // (y0 + y2) + (y1 + y3)
        }
    }
}

// Exclusive scan of one sequence's segment maps in flow order -> the state entering each segment.
// grid (B * D, 4): blockIdx.y = direction; 256 threads walk the NS segments 256 at a time (coalesced), carrying the running map.
__global__ void __launch_bounds__(256) ss2d_carry_kernel(const Ss2dFusedArgs p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float2 wtot[8];
    const int k = blockIdx.y;
    const int64_t bd = blockIdx.x;
    const int64_t NSr = (int64_t)p.H * p.NTW, NSc = (int64_t)p.W * p.NTH;
    const int64_t NS = (k & 1) ? NSc : NSr;
    const int64_t dir_off = seg_offset(p, k);
    const float2* agg = p.agg + dir_off + bd * NS;
    float* hin = p.hin + dir_off + bd * NS;
    const bool rev = k >= 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float carry = 0.f;                                   // state entering the current block of 256 segments
    for (int64_t f0 = 0; f0 < NS; f0 += 256) {
        const int64_t f = f0 + threadIdx.x;              // flow index
        const int64_t s = rev ? NS - 1 - f : f;          // memory index
        float P = 1.f, V = 0.f;
        if (f < NS) {
            const float2 a = agg[s];
            P = a.x;
            V = a.y;
        }
        warp_scan_fwd(P, V, lane);                       // inclusive over the warp, flow order = lane order
        if (lane == 31) wtot[warp] = make_float2(P, V);
        __syncthreads();
        // state entering this warp's first segment: carry pushed through the preceding warps of the block
        float hw = carry;
        for (int w2 = 0; w2 < warp; ++w2) hw = fmaf(wtot[w2].x, hw, wtot[w2].y);
        // exclusive value of this lane: map of lanes 0 .. lane-1 applied to hw
        float Pe = __shfl_up_sync(FULL, P, 1), Ve = __shfl_up_sync(FULL, V, 1);
        if (lane == 0) {
            Pe = 1.f;
            Ve = 0.f;
        }
        if (f < NS) hin[s] = fmaf(Pe, hw, Ve);
        float c = carry;
        for (int w2 = 0; w2 < 8; ++w2) c = fmaf(wtot[w2].x, c, wtot[w2].y);
        carry = c;
        __syncthreads();
    }
}

template <int R>
static int launch_fused(const Ss2dFusedArgs& a, cudaStream_t stream) {
    constexpr int CX = R + 2;
    const size_t sm1 = (size_t)(DB * TS * PITCH + 2 * CX * TS * PITCH) * 4;
    const size_t sm3 = sm1 + (size_t)DB * TS * PITCH * 4;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ss2d_tile_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(ss2d_tile_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
        if (e != cudaSuccess) return (int)e;
        attr_done[dev] = true;
    }
    const int nDb = (a.D + DB - 1) / DB;
    const dim3 grid(a.NTW, a.NTH, a.B * nDb);
    if ((int64_t)a.B * nDb > 65535 || a.NTH > 65535) return BEM_ERR_UNSUPPORTED;
    launch_pdl(ss2d_tile_kernel<R, false>, grid, dim3(THREADS), sm1, stream, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    launch_pdl(ss2d_carry_kernel, dim3(a.B * a.D, 4), dim3(256), 0, stream, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    launch_pdl(ss2d_tile_kernel<R, true>, grid, dim3(THREADS), sm3, stream, a);
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
        case 3: return launch_fused<3>(a, stream);
        case 5: return launch_fused<5>(a, stream);
        case 10: return launch_fused<10>(a, stream);
        default: return BEM_ERR_UNSUPPORTED;
    }
}

}  // namespace bem
